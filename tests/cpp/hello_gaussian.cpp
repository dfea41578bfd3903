// hello_gaussian.cpp — demo/HelloGaussian/main.cpp (reference, :13-58) on the Vulkan-free drop-in:
// a cube of random Gaussians plus one big white Gaussian at the origin, SH degree 0, orbit camera at
// (theta 0.785, phi 0.9, radius 8), one frame, written as a binary PPM. Prints the pair count and a checksum.
//   g++ -std=c++20 -O2 -ffp-contract=off -Iinclude tests/cpp/hello_gaussian.cpp -Ltorpedo_b200/lib -ltpdcu -o hello
#include <torpedo_b200/GaussianEngine.hpp>

#include <cstdio>
#include <fstream>
#include <vector>

int main(int argc, char** argv) {
    constexpr uint32_t width = 1280, height = 720;
    try {
        auto engine = tpd::GaussianEngine{ width, height };
        auto camera = tpd::PerspectiveCamera{ width, height };

        // Generate random points within a cube of size 10 centered at the origin
        auto points = tpd::GaussianPoint::random(8192, 10.f, {}, 0.005f, 0.2f, 0.1f, 1.0f, /*seed*/ 1);

        // A big, white, uniform Gaussian at the center of the scene
        auto gaussian = tpd::GaussianPoint{
            .position = { 0.f, 0.f, 0.f },
            .opacity = 1.f,
            .quaternion = { 0.f, 0.f, 0.f, 1.f },
            .scale = { 2.f, 2.f, 2.f, 1.0f },
            .sh = tpd::utils::rgb2sh(1.0f, 1.0f, 1.0f),
        };

        auto scene = tpd::Scene{};
        scene.add(tpd::ent::group(points));
        scene.add(std::move(gaussian));

        auto settings = tpd::GaussianEngine::Settings::getDefault();
        settings.sphericalHarmonicsDegree = 0;

        engine.compile(scene, settings);
        points.clear();  // all data has been transferred to the GPU

        camera.lookAt(tpd::math::to_cartesian(0.785f, 0.9f, 8.f), { 0.f, 0.f, 0.f }, { 0.f, 0.f, 1.f });
        engine.rasterFrame(camera);

        std::vector<unsigned char> rgba(size_t(width) * height * 4);
        engine.draw(rgba.data(), size_t(width) * 4);

        uint32_t pairs = 0, visible = 0;
        tpdcu_get_counts(engine.handle(), &pairs, &visible);
        unsigned long long checksum = 0;
        for (size_t i = 0; i < rgba.size(); i += 4) checksum += rgba[i] + rgba[i + 1] + rgba[i + 2];
        std::printf("pairs %u visible %u checksum %llu\n", pairs, visible, checksum);

        if (argc > 1) {
            std::ofstream ppm(argv[1], std::ios::binary);
            ppm << "P6\n" << width << " " << height << "\n255\n";
            for (size_t i = 0; i < rgba.size(); i += 4) ppm.write(reinterpret_cast<const char*>(&rgba[i]), 3);
        }
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "hello_gaussian: %s\n", e.what());
        return 1;
    }
}
