// tex_weight_probe.cu -- how finely does the texture unit resolve the filter position of a LINEAR tex3D fetch?
// A 4x4x4 RGBA8 volume whose red channel is 0 in texel x = 1 and 255 in texel x = 2 (constant along y, z); 2^16 fetches sweep x across that
// cell at y, z texel centres (and, second pass, with y and z off-centre); the distinct filtered values and the width of their steps
// give the weight resolution.  Third pass: the same sweep at a large normalized coordinate (x + 24, WRAP), as the march uses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tex_weight_probe tools/probes/tex_weight_probe.cu && /tmp/tex_weight_probe
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

__global__ void sweep(cudaTextureObject_t tex, float x0, float x1, int n, float y, float z, float* out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = x0 + (x1 - x0) * ((float)i / (float)n);
    out[i] = tex3D<float4>(tex, x, y, z).x;
}

int main()
{
    const int W = 4;
    std::vector<unsigned char> vol(W * W * W * 4, 0);
    for (int z = 0; z < W; ++z)
        for (int y = 0; y < W; ++y) vol[((z * W + y) * W + 2) * 4] = 255;
    cudaArray_t arr;
    cudaChannelFormatDesc fd = cudaCreateChannelDesc<uchar4>();
    cudaMalloc3DArray(&arr, &fd, make_cudaExtent(W, W, W));
    cudaMemcpy3DParms cp = {};
    cp.srcPtr = make_cudaPitchedPtr(vol.data(), W * 4, W, W);
    cp.dstArray = arr;
    cp.extent = make_cudaExtent(W, W, W);
    cp.kind = cudaMemcpyHostToDevice;
    cudaMemcpy3D(&cp);
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = arr;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeWrap;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 1;
    cudaTextureObject_t tex;
    cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    const int n = 1 << 16;
    float* d;
    cudaMalloc(&d, n * sizeof(float));
    std::vector<float> h(n);
    // texel centres: (i + 0.5) / W; the cell between texel 1 and texel 2 spans x in [1.5 / W, 2.5 / W]
    const float cases[3][3] = { { 0.0f, 1.5f / W, 1.5f / W }, { 0.0f, 1.8f / W, 1.3f / W }, { 24.0f, 1.5f / W, 1.5f / W } };
    for (int c = 0; c < 3; ++c) {
        sweep<<<n / 256, 256>>>(tex, cases[c][0] + 1.5f / W, cases[c][0] + 2.5f / W, n, cases[c][1], cases[c][2], d);
        cudaMemcpy(h.data(), d, n * sizeof(float), cudaMemcpyDeviceToHost);
        int levels = 1, first = -1, last = -1;
        for (int i = 1; i < n; ++i)
            if (h[i] != h[i - 1]) { ++levels; if (first < 0) first = i; last = i; }
        printf("case %d (x offset %g, y %.3f, z %.3f): %d distinct levels over one cell, min %.6f max %.6f, first change at %.5f of the cell, last at %.5f"
               " => weight step 1/%d\n", c, cases[c][0], cases[c][1] * W, cases[c][2] * W, levels, h[0], h[n - 1], (double)first / n, (double)last / n, levels - 1);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
