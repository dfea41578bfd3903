// adapters/vulkan_interop.cpp — the Vulkan side of the frame hand-off.
//
// NOT BUILT OR RUN IN THIS IMAGE: there are no Vulkan headers, loader or ICD here or on the GPU boxes. What is checked: the file
// compiles warning-free against declarations of the Vulkan symbols it uses (tests/cpp/vulkan_stub, written from the specification;
// tests/test_adapter_syntax.py) and its object references tpdcu_bind_output_fd of include/tpdcu.h. The CUDA half of every
// call below is executed by tests/test_external_memory_gpu.py (an opaque POSIX fd exported by the CUDA driver's VMM API is
// imported through tpdcu_bind_output_fd and a frame is rendered into it); this file is the other half, written against
// plain vulkan.h so that it does not depend on the Vulkan-Hpp version torpedo pins. Build it inside torpedo with
//   target_sources(torpedo_volumetric PRIVATE ${TORPEDO_B200_DIR}/adapters/vulkan_interop.cpp)
//
// What it replaces in the reference:
//   rendering/src/Engine.cpp:42-49            getDeviceExtensions(): + VK_KHR_external_memory_fd (+ semaphore_fd)
//   foundation/src/VmaUsage.cpp:4-62          target allocation: a dedicated, EXPORTABLE VkBuffer instead of a VkImage
//   volumetric/src/GaussianEngine.cpp:315-333 createRenderTargets
//   volumetric/src/GaussianEngine.cpp:714-762 draw(): recordTargetCopy image -> swap image becomes buffer -> swap image
//   volumetric/src/GaussianEngine.cpp:865-875 recordTargetCopy
#include <vulkan/vulkan.h>

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include <cuda_runtime_api.h>

#include "../include/tpdcu.h"

namespace tpd::interop {

// Device extensions GaussianEngine::getDeviceExtensions() has to add (VK_KHR_external_memory itself is core in 1.1).
inline std::vector<const char*> requiredDeviceExtensions(bool withSemaphore) {
    std::vector<const char*> ext{ VK_KHR_EXTERNAL_MEMORY_FD_EXTENSION_NAME };
    if (withSemaphore) ext.push_back(VK_KHR_EXTERNAL_SEMAPHORE_FD_EXTENSION_NAME);
    return ext;
}

// The CUDA ordinal of the GPU Vulkan renders on: VkPhysicalDeviceIDProperties::deviceUUID == cudaDeviceProp::uuid.
inline int cudaDeviceOf(VkPhysicalDevice physical) {
    VkPhysicalDeviceIDProperties id{ VK_STRUCTURE_TYPE_PHYSICAL_DEVICE_ID_PROPERTIES };
    VkPhysicalDeviceProperties2 props{ VK_STRUCTURE_TYPE_PHYSICAL_DEVICE_PROPERTIES_2, &id };
    vkGetPhysicalDeviceProperties2(physical, &props);
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) throw std::runtime_error("tpd::interop - no CUDA device");
    for (int d = 0; d < count; ++d) {
        cudaDeviceProp p{};
        if (cudaGetDeviceProperties(&p, d) == cudaSuccess && std::memcmp(p.uuid.bytes, id.deviceUUID, VK_UUID_SIZE) == 0) return d;
    }
    throw std::runtime_error("tpd::interop - the Vulkan physical device is not a CUDA device");
}

// One linear RGBA8 frame the CUDA rasterizer writes and the graphics queue copies into the swap image: the per-frame render
// target of the reference (GaussianEngine::Frame::outputImage, GaussianEngine.h:104-117) as an exportable buffer.
class PresentTarget {
public:
    PresentTarget(VkPhysicalDevice physical, VkDevice device, uint32_t width, uint32_t height)
        : _device{ device }, _width{ width }, _height{ height }, _bytes{ VkDeviceSize(width) * height * 4 } {
        VkExternalMemoryBufferCreateInfo external{ VK_STRUCTURE_TYPE_EXTERNAL_MEMORY_BUFFER_CREATE_INFO };
        external.handleTypes = VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT;
        VkBufferCreateInfo info{ VK_STRUCTURE_TYPE_BUFFER_CREATE_INFO, &external };
        info.size = _bytes;
        info.usage = VK_BUFFER_USAGE_TRANSFER_SRC_BIT;
        info.sharingMode = VK_SHARING_MODE_EXCLUSIVE;
        vk(vkCreateBuffer(device, &info, nullptr, &_buffer), "vkCreateBuffer");

        VkMemoryRequirements req{};
        vkGetBufferMemoryRequirements(device, _buffer, &req);
        VkPhysicalDeviceMemoryProperties mem{};
        vkGetPhysicalDeviceMemoryProperties(physical, &mem);
        uint32_t type = UINT32_MAX;
        for (uint32_t i = 0; i < mem.memoryTypeCount; ++i)
            if ((req.memoryTypeBits & (1u << i)) && (mem.memoryTypes[i].propertyFlags & VK_MEMORY_PROPERTY_DEVICE_LOCAL_BIT)) { type = i; break; }
        if (type == UINT32_MAX) throw std::runtime_error("tpd::interop - no device-local memory type for the present target");

        // dedicated + exportable: what VMA's DEDICATED_MEMORY_BIT targets are in the reference (VmaUsage.cpp:28-42); the
        // importer passes cudaExternalMemoryDedicated for it (tpdcu_bind_output_fd)
        VkMemoryDedicatedAllocateInfo dedicated{ VK_STRUCTURE_TYPE_MEMORY_DEDICATED_ALLOCATE_INFO };
        dedicated.buffer = _buffer;
        VkExportMemoryAllocateInfo exported{ VK_STRUCTURE_TYPE_EXPORT_MEMORY_ALLOCATE_INFO, &dedicated };
        exported.handleTypes = VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT;
        VkMemoryAllocateInfo alloc{ VK_STRUCTURE_TYPE_MEMORY_ALLOCATE_INFO, &exported };
        alloc.allocationSize = req.size;
        alloc.memoryTypeIndex = type;
        vk(vkAllocateMemory(device, &alloc, nullptr, &_memory), "vkAllocateMemory");
        vk(vkBindBufferMemory(device, _buffer, _memory, 0), "vkBindBufferMemory");
        _allocationBytes = req.size;
    }

    PresentTarget(const PresentTarget&) = delete;
    PresentTarget& operator=(const PresentTarget&) = delete;

    ~PresentTarget() {
        if (_buffer) vkDestroyBuffer(_device, _buffer, nullptr);
        if (_memory) vkFreeMemory(_device, _memory, nullptr);
    }

    // Hand the memory to the rasterizer: from now on tpdcu_raster renders into this buffer. A successful import owns the fd.
    void bind(tpdcu_ctx* cuda) {
        auto getFd = reinterpret_cast<PFN_vkGetMemoryFdKHR>(vkGetDeviceProcAddr(_device, "vkGetMemoryFdKHR"));
        if (!getFd) throw std::runtime_error("tpd::interop - VK_KHR_external_memory_fd is not enabled on the device");
        VkMemoryGetFdInfoKHR info{ VK_STRUCTURE_TYPE_MEMORY_GET_FD_INFO_KHR };
        info.memory = _memory;
        info.handleType = VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT;
        int fd = -1;
        vk(getFd(_device, &info, &fd), "vkGetMemoryFdKHR");
        if (tpdcu_bind_output_fd(cuda, fd, static_cast<size_t>(_allocationBytes)) != TPDCU_OK)
            throw std::runtime_error(std::string("tpd::interop - ") + tpdcu_last_error());
    }

    // GaussianEngine::draw (GaussianEngine.cpp:714-762) with recordTargetCopy (:865-875) turned into a buffer -> image copy.
    // The caller has finished the CUDA frame (tpdcu_finish, or a wait on an imported timeline semaphore)
    // and keeps the reference's layout transitions around this call: swapImage is in TRANSFER_DST_OPTIMAL here.
    void recordCopyToSwapImage(VkCommandBuffer cmd, VkImage swapImage) const {
        VkBufferMemoryBarrier acquire{ VK_STRUCTURE_TYPE_BUFFER_MEMORY_BARRIER };
        acquire.srcAccessMask = 0;
        acquire.dstAccessMask = VK_ACCESS_TRANSFER_READ_BIT;
        acquire.srcQueueFamilyIndex = VK_QUEUE_FAMILY_EXTERNAL;   // written outside Vulkan: acquire from the external owner
        acquire.dstQueueFamilyIndex = VK_QUEUE_FAMILY_IGNORED;
        acquire.buffer = _buffer;
        acquire.size = _bytes;
        vkCmdPipelineBarrier(cmd, VK_PIPELINE_STAGE_TOP_OF_PIPE_BIT, VK_PIPELINE_STAGE_TRANSFER_BIT, 0, 0, nullptr, 1, &acquire, 0, nullptr);
        VkBufferImageCopy region{};
        region.bufferRowLength = _width;      // tightly packed rows: tpdcu_bind_output_fd sets pitch = width * 4
        region.bufferImageHeight = _height;
        region.imageSubresource = { VK_IMAGE_ASPECT_COLOR_BIT, 0, 0, 1 };
        region.imageExtent = { _width, _height, 1 };
        vkCmdCopyBufferToImage(cmd, _buffer, swapImage, VK_IMAGE_LAYOUT_TRANSFER_DST_OPTIMAL, 1, &region);
    }

    VkBuffer buffer() const { return _buffer; }

private:
    static void vk(VkResult r, const char* what) {
        if (r != VK_SUCCESS) throw std::runtime_error(std::string("tpd::interop - ") + what + " failed: " + std::to_string(int(r)));
    }
    VkDevice _device;
    uint32_t _width, _height;
    VkDeviceSize _bytes, _allocationBytes = 0;
    VkBuffer _buffer = VK_NULL_HANDLE;
    VkDeviceMemory _memory = VK_NULL_HANDLE;
};

}  // namespace tpd::interop
