// Instantiates every entry point of adapters/vulkan_interop.cpp so that a compile to an object file type-checks their bodies
// against tests/cpp/vulkan_stub (declarations only) and include/tpdcu.h; tests/test_adapter_syntax.py then reads the object's
// undefined symbols: the Vulkan entry points and tpdcu_bind_output_fd the adapter binds to. Nothing here is linked or run.
#include "../../adapters/vulkan_interop.cpp"

int adapter_use(VkPhysicalDevice physical, VkDevice device, VkCommandBuffer cmd, VkImage swapImage, tpdcu_ctx* cuda) {
    tpd::interop::PresentTarget target(physical, device, 1920, 1080);
    target.bind(cuda);
    target.recordCopyToSwapImage(cmd, swapImage);
    return tpd::interop::cudaDeviceOf(physical) + int(tpd::interop::requiredDeviceExtensions(true).size()) + (target.buffer() != nullptr);
}
