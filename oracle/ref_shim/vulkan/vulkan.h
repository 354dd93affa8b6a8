/* TEST INFRASTRUCTURE -- a stand-in for <vulkan/vulkan.h>, just enough declarations for the reference's host-side
 * uniform producers (camera.cpp, Scene.cpp, Sky.cpp) and its texture loader (ImageLoadingUtility.cpp) to compile
 * unmodified where they lie under /root/reference.
 * There is no Vulkan loader in this image; oracle/ref_harness.cpp backs "device memory" with malloc.  Nothing here
 * is taken from the Khronos headers beyond the public names the reference spells. */
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <assert.h>
#include <string>
#include <cmath>

#define MT_VK_HANDLE(name) typedef struct name##_T* name
MT_VK_HANDLE(VkInstance); MT_VK_HANDLE(VkPhysicalDevice); MT_VK_HANDLE(VkDevice); MT_VK_HANDLE(VkQueue);
MT_VK_HANDLE(VkSemaphore); MT_VK_HANDLE(VkCommandBuffer); MT_VK_HANDLE(VkCommandPool); MT_VK_HANDLE(VkBuffer);
MT_VK_HANDLE(VkDeviceMemory); MT_VK_HANDLE(VkImage); MT_VK_HANDLE(VkImageView); MT_VK_HANDLE(VkSampler);
MT_VK_HANDLE(VkSurfaceKHR); MT_VK_HANDLE(VkSwapchainKHR); MT_VK_HANDLE(VkDebugReportCallbackEXT);
MT_VK_HANDLE(VkDescriptorSet); MT_VK_HANDLE(VkDescriptorSetLayout); MT_VK_HANDLE(VkDescriptorPool);
MT_VK_HANDLE(VkPipeline); MT_VK_HANDLE(VkPipelineLayout); MT_VK_HANDLE(VkShaderModule); MT_VK_HANDLE(VkFramebuffer);
MT_VK_HANDLE(VkRenderPass);
#define VK_NULL_HANDLE 0

typedef uint32_t VkFlags;
typedef uint32_t VkBool32;
typedef uint64_t VkDeviceSize;
typedef VkFlags VkMemoryPropertyFlags, VkBufferUsageFlags, VkImageUsageFlags, VkMemoryMapFlags, VkImageAspectFlags;
typedef int VkResult, VkFormat, VkPresentModeKHR, VkImageTiling, VkImageLayout, VkSamplerAddressMode, VkVertexInputRate;
typedef int VkStructureType, VkImageType, VkSharingMode, VkSampleCountFlagBits;
struct VkAllocationCallbacks;

enum {
    VK_SUCCESS = 0,
    VK_BUFFER_USAGE_TRANSFER_SRC_BIT = 0x1, VK_BUFFER_USAGE_UNIFORM_BUFFER_BIT = 0x10,
    VK_MEMORY_PROPERTY_DEVICE_LOCAL_BIT = 0x1, VK_MEMORY_PROPERTY_HOST_VISIBLE_BIT = 0x2, VK_MEMORY_PROPERTY_HOST_COHERENT_BIT = 0x4,
    VK_IMAGE_USAGE_TRANSFER_DST_BIT = 0x2, VK_IMAGE_USAGE_SAMPLED_BIT = 0x4,
    VK_FORMAT_R8G8B8A8_UNORM = 37, VK_FORMAT_R32G32_SFLOAT = 103, VK_FORMAT_R32G32B32_SFLOAT = 106, VK_FORMAT_R32G32B32A32_SFLOAT = 109,
    VK_IMAGE_TILING_OPTIMAL = 0, VK_SAMPLER_ADDRESS_MODE_REPEAT = 0, VK_VERTEX_INPUT_RATE_VERTEX = 0,
    VK_IMAGE_LAYOUT_UNDEFINED = 0, VK_IMAGE_LAYOUT_SHADER_READ_ONLY_OPTIMAL = 5, VK_IMAGE_LAYOUT_TRANSFER_DST_OPTIMAL = 7,
    VK_STRUCTURE_TYPE_MEMORY_ALLOCATE_INFO = 5, VK_STRUCTURE_TYPE_IMAGE_CREATE_INFO = 14,
    VK_IMAGE_TYPE_3D = 2, VK_SHARING_MODE_EXCLUSIVE = 0, VK_SAMPLE_COUNT_1_BIT = 1,
};

struct VkExtent2D { uint32_t width, height; };
struct VkSurfaceFormatKHR { VkFormat format; int colorSpace; };
struct VkSurfaceCapabilitiesKHR { uint32_t minImageCount, maxImageCount; VkExtent2D currentExtent, minImageExtent, maxImageExtent; };
struct VkPhysicalDeviceMemoryProperties { uint32_t memoryTypeCount; };
struct VkVertexInputBindingDescription { uint32_t binding, stride; VkVertexInputRate inputRate; };
struct VkVertexInputAttributeDescription { uint32_t location, binding; VkFormat format; uint32_t offset; };

struct VkExtent3D { uint32_t width, height, depth; };
struct VkImageCreateInfo {
    VkStructureType sType; const void* pNext; VkFlags flags; VkImageType imageType; VkFormat format; VkExtent3D extent;
    uint32_t mipLevels, arrayLayers; VkSampleCountFlagBits samples; VkImageTiling tiling; VkImageUsageFlags usage;
    VkSharingMode sharingMode; VkImageLayout initialLayout;
};
struct VkMemoryRequirements { VkDeviceSize size, alignment; uint32_t memoryTypeBits; };
struct VkMemoryAllocateInfo { VkStructureType sType; const void* pNext; VkDeviceSize allocationSize; uint32_t memoryTypeIndex; };

VkResult vkCreateImage(VkDevice, const VkImageCreateInfo*, const VkAllocationCallbacks*, VkImage*);
void vkGetImageMemoryRequirements(VkDevice, VkImage, VkMemoryRequirements*);
VkResult vkAllocateMemory(VkDevice, const VkMemoryAllocateInfo*, const VkAllocationCallbacks*, VkDeviceMemory*);
VkResult vkBindImageMemory(VkDevice, VkImage, VkDeviceMemory, VkDeviceSize);
VkResult vkMapMemory(VkDevice, VkDeviceMemory, VkDeviceSize offset, VkDeviceSize size, VkMemoryMapFlags, void** ppData);
void vkUnmapMemory(VkDevice, VkDeviceMemory);
void vkDestroyBuffer(VkDevice, VkBuffer, const VkAllocationCallbacks*);
void vkFreeMemory(VkDevice, VkDeviceMemory, const VkAllocationCallbacks*);
