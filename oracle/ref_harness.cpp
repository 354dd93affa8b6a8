// TEST INFRASTRUCTURE -- not part of the product, never linked into libmeteoros_b200.so.
//
// Runs the reference's OWN host-side code for the inputs of the hot path -- the uniform producers camera.cpp,
// Scene.cpp, Sky.cpp and the texture loader ImageLoadingUtility.cpp (with the vendored stb_image.h), compiled
// unmodified from /root/reference/src/CloudScapes with the reference's vendored glm -- and hands the bytes they would
// give the GPU back through a plain C interface.  This pins SURVEY.md 8(a) A0 (CameraUBO, Time, SunAndSky; N2 mirrors
// them in mt_scene.cpp / scene.py) and the texture memory layout (N3, mt_assets.cpp) against the real code instead of
// a second restatement.
//
// These files talk to Vulkan only to create a buffer / image, map it and memcpy into it (plus layout transitions
// and the buffer -> image copy).  oracle/ref_shim/vulkan/vulkan.h declares those few entry points; this file
// implements them over malloc, so the bytes the reference would hand to the GPU land in memory we can read.  The
// classes those files merely mention (Model, Texture2D, Texture3D, VulkanDevice) get do-nothing definitions.  Built by
// `make -C oracle ref` into oracle/_ref/libmeteoros_ref.so, only where /root/reference exists;
// tests/golden/make_ref_uniforms.py turns its output into the committed fixture tests/golden/ref_uniforms.npz.
#include "camera.h"
#include "Scene.h"
#include "Sky.h"
#include "ImageLoadingUtility.h"

#include <cstdlib>
#include <cstring>

// ---- "Vulkan" over malloc ---------------------------------------------------------------------------------------
struct Block { size_t size; unsigned char* bytes; };

void BufferUtils::CreateBuffer(VulkanDevice*, VkBufferUsageFlags, VkDeviceSize size, VkMemoryPropertyFlags, VkBuffer& buffer,
                               VkDeviceMemory& bufferMemory)
{
    Block* b = new Block{ (size_t)size, (unsigned char*)calloc(1, (size_t)size) };
    buffer = reinterpret_cast<VkBuffer>(b);
    bufferMemory = reinterpret_cast<VkDeviceMemory>(b);
}
VkResult vkMapMemory(VkDevice, VkDeviceMemory m, VkDeviceSize offset, VkDeviceSize, VkMemoryMapFlags, void** out)
{
    *out = reinterpret_cast<Block*>(m)->bytes + offset;
    return VK_SUCCESS;
}
void vkUnmapMemory(VkDevice, VkDeviceMemory) {}
void vkDestroyBuffer(VkDevice, VkBuffer, const VkAllocationCallbacks*) {}
void vkFreeMemory(VkDevice, VkDeviceMemory m, const VkAllocationCallbacks*)
{
    Block* b = reinterpret_cast<Block*>(m);
    free(b->bytes);
    delete b;
}
VkDevice VulkanDevice::GetVkDevice() { return nullptr; }
VulkanInstance* VulkanDevice::GetInstance() { return reinterpret_cast<VulkanInstance*>(sizeof(void*)); }
uint32_t VulkanInstance::GetMemoryTypeIndex(uint32_t, VkMemoryPropertyFlags) const { return 0; }

// images: a malloc block of width * height * depth RGBA8 texels, filled by the "buffer -> image copy"
static VkImage make_image(size_t texels)
{
    return reinterpret_cast<VkImage>(new Block{ texels * 4, (unsigned char*)calloc(texels, 4) });
}
VkResult vkCreateImage(VkDevice, const VkImageCreateInfo* ci, const VkAllocationCallbacks*, VkImage* image)
{
    *image = make_image((size_t)ci->extent.width * ci->extent.height * ci->extent.depth);
    return VK_SUCCESS;
}
void vkGetImageMemoryRequirements(VkDevice, VkImage image, VkMemoryRequirements* r)
{
    r->size = reinterpret_cast<Block*>(image)->size; r->alignment = 16; r->memoryTypeBits = 1;
}
VkResult vkAllocateMemory(VkDevice, const VkMemoryAllocateInfo*, const VkAllocationCallbacks*, VkDeviceMemory* m) { *m = nullptr; return VK_SUCCESS; }
VkResult vkBindImageMemory(VkDevice, VkImage, VkDeviceMemory, VkDeviceSize) { return VK_SUCCESS; }
void Image::createImage(VulkanDevice*, uint32_t width, uint32_t height, VkFormat, VkImageTiling, VkImageUsageFlags, VkMemoryPropertyFlags,
                        VkImage& image, VkDeviceMemory& imageMemory)
{
    image = make_image((size_t)width * height);
    imageMemory = nullptr;
}
void Image::transitionImageLayout(VulkanDevice*, VkCommandPool, VkImage&, VkFormat, VkImageLayout, VkImageLayout) {}
static void copy_to_image(VkBuffer buffer, VkImage image, size_t texels)   // vkCmdCopyBufferToImage, tightly packed (Image.cpp:9-79)
{
    Block* src = reinterpret_cast<Block*>(buffer);
    Block* dst = reinterpret_cast<Block*>(image);
    if (texels * 4 > src->size || texels * 4 > dst->size) abort();
    memcpy(dst->bytes, src->bytes, texels * 4);
}
void Image::copyBufferToImage(VulkanDevice*, VkCommandPool, VkBuffer buffer, VkImage& image, uint32_t width, uint32_t height)
{
    copy_to_image(buffer, image, (size_t)width * height);
}
void Image::copyBufferToImage3D(VulkanDevice*, VkCommandPool, VkBuffer buffer, VkImage& image, uint32_t width, uint32_t height, uint32_t depth)
{
    copy_to_image(buffer, image, (size_t)width * height * depth);
}

// mentioned by Scene.cpp / Sky.cpp, never reached from the calls below
Model::Model(VulkanDevice*, VkCommandPool, const std::string, const std::string) {}
Model::~Model() {}
glm::mat4 Model::GetModelMatrix() const { return glm::mat4(1.0f); }
void Model::SetModelBuffer(glm::mat4&) {}
Texture2D::Texture2D(VulkanDevice*, uint32_t, uint32_t, VkFormat) {}
Texture2D::~Texture2D() {}
void Texture2D::createTextureFromFile(VkDevice, VkCommandPool, const std::string, int, VkImageTiling, VkImageUsageFlags,
                                      VkMemoryPropertyFlags, VkSamplerAddressMode, float) {}
Texture3D::Texture3D(VulkanDevice*, uint32_t, uint32_t, uint32_t, VkFormat) {}
Texture3D::~Texture3D() {}
void Texture3D::create3DTextureFromMany2DTextures(VkDevice, VkCommandPool, const std::string, const std::string, const std::string,
                                                  int, int) {}

static VulkanDevice* fake_device() { return reinterpret_cast<VulkanDevice*>(sizeof(void*)); }  // only ever passed along
static const void* mapped_bytes(VkBuffer b) { return reinterpret_cast<Block*>(b)->bytes; }

extern "C" {

enum { MTREF_ROTATE_UP = 0, MTREF_ROTATE_RIGHT, MTREF_ALONG_LOOK, MTREF_ALONG_RIGHT, MTREF_ALONG_UP };

// Camera(...) as main.cpp:157-158 builds it, then `n_ops` control calls (main.cpp:60-110), each followed by what the
// frame loop does: UpdateBuffer() + CopyToGPUMemory() (main.cpp:176-177).  `out` receives (n_ops + 1) * 152 bytes:
// the mapped uniform buffer after construction and after every op.
int mtref_camera(const float eye[3], const float ref[3], int width, int height, float fovy, float near_clip, float far_clip,
                 int n_ops, const int* ops, const float* args, unsigned char* out)
{
    static_assert(sizeof(CameraUBO) == 152, "CameraUBO layout");
    Camera cam(fake_device(), glm::vec3(eye[0], eye[1], eye[2]), glm::vec3(ref[0], ref[1], ref[2]), width, height, fovy,
               width / (float)height, near_clip, far_clip);
    memcpy(out, mapped_bytes(cam.GetBuffer()), sizeof(CameraUBO));
    for (int i = 0; i < n_ops; ++i) {
        switch (ops[i]) {
        case MTREF_ROTATE_UP: cam.RotateAboutUp(args[i]); break;
        case MTREF_ROTATE_RIGHT: cam.RotateAboutRight(args[i]); break;
        case MTREF_ALONG_LOOK: cam.TranslateAlongLook(args[i]); break;
        case MTREF_ALONG_RIGHT: cam.TranslateAlongRight(args[i]); break;
        case MTREF_ALONG_UP: cam.TranslateAlongUp(args[i]); break;
        default: return -1;
        }
        cam.UpdateBuffer();
        cam.CopyToGPUMemory();
        memcpy(out + (size_t)(i + 1) * sizeof(CameraUBO), mapped_bytes(cam.GetBuffer()), sizeof(CameraUBO));
    }
    return 0;
}

// Scene(device) -> InitializeTime() (Scene.cpp:86-119), then `n_updates` x UpdateTime() (Scene.cpp:65-85).
// `out` receives (n_updates + 1) * 76 bytes.  Bytes 64..71 (delta / total time) come from the wall clock.
int mtref_time(int n_updates, unsigned char* out)
{
    static_assert(sizeof(Time) == 76, "Time layout");
    Scene scene(fake_device());
    memcpy(out, mapped_bytes(scene.GetTimeBuffer()), sizeof(Time));
    for (int i = 0; i < n_updates; ++i) {
        scene.UpdateTime();
        memcpy(out + (size_t)(i + 1) * sizeof(Time), mapped_bytes(scene.GetTimeBuffer()), sizeof(Time));
    }
    return 0;
}

float mtref_halton(int index, int base)
{
    Scene scene(fake_device());
    return scene.HaltonSequenceAt(index, base);
}

// Sky(device) + UpdateSunAndSky() (Sky.cpp:64-74): 52 bytes.
int mtref_sun_and_sky(unsigned char* out)
{
    static_assert(sizeof(SunAndSky) == 52, "SunAndSky layout");
    Sky* sky = new Sky(fake_device(), nullptr);
    sky->weatherMapTexture = nullptr; sky->cloudMotionTexture = nullptr;   // ~Sky deletes them; CreateCloudResources never ran
    sky->cloudBaseShapeTexture = nullptr; sky->cloudDetailsTexture = nullptr;
    sky->UpdateSunAndSky();
    memcpy(out, mapped_bytes(sky->GetSunAndSkyBuffer()), sizeof(SunAndSky));
    delete sky;
    return 0;
}

// Sky::CreateCloudResources -> Texture3D::create3DTextureFromMany2DTextures -> ImageLoadingUtility (..cpp:75-139):
// `depth` slices "<folder><base>(z+1)<ext>" decoded by stb_image and packed [z][y][x][rgba].  `out` receives the image.
int mtref_load_volume(const char* folder, const char* base, const char* ext, int width, int height, int depth, unsigned char* out)
{
    VkImage image = nullptr;
    VkDeviceMemory memory = nullptr;
    try {
        ImageLoadingUtility::create3DTextureFromMany2DTextures(fake_device(), nullptr, nullptr, folder, base, ext, image, memory,
                                                               VK_FORMAT_R8G8B8A8_UNORM, width, height, depth, depth, 4);
    } catch (const std::exception&) {
        return -1;
    }
    memcpy(out, reinterpret_cast<Block*>(image)->bytes, (size_t)width * height * depth * 4);
    return 0;
}

// Texture2D::createTextureFromFile -> ImageLoadingUtility::loadImageFromFile (..cpp:9-73): stbi_load(path, 4 channels).
// `out` receives width * height * 4 bytes; the size must be known to the caller (as it is to Sky.cpp:47-57).
int mtref_load_image(const char* path, int width, int height, unsigned char* out)
{
    VkImage image = nullptr;
    VkDeviceMemory memory = nullptr;
    VkCommandPool pool = nullptr;
    try {
        ImageLoadingUtility::loadImageFromFile(fake_device(), pool, path, image, memory, VK_FORMAT_R8G8B8A8_UNORM, VK_IMAGE_TILING_OPTIMAL,
                                               VK_IMAGE_USAGE_TRANSFER_DST_BIT | VK_IMAGE_USAGE_SAMPLED_BIT, VK_MEMORY_PROPERTY_DEVICE_LOCAL_BIT);
    } catch (const std::exception&) {
        return -1;
    }
    if (reinterpret_cast<Block*>(image)->size != (size_t)width * height * 4) return -2;
    memcpy(out, reinterpret_cast<Block*>(image)->bytes, (size_t)width * height * 4);
    return 0;
}

}  // extern "C"
