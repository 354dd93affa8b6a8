// mt_launch.h -- launchers of the pass kernels (implemented in the .cu files, called by mt_context.cu).
#pragma once

#include <cuda_runtime.h>

#include "mt_params.h"

cudaError_t mt_launch_cloud(const CloudParams& P, cudaStream_t stream);
cudaError_t mt_launch_tile_forward(const void* src, void* dst, int W, int H, int bytesPerPixel, const RowTiles& rows, unsigned* tileDone,
                                   int ctas, cudaStream_t stream);
cudaError_t mt_launch_cloud_sixteenth_split(const CloudParams& P, cudaStream_t stream, int* launches);
cudaError_t mt_launch_cloud_sixteenth_fused(const CloudParams& P, cudaStream_t stream);
cudaError_t mt_launch_build_quads(const uint32_t* texels, int w, int h, int d, void* quads, cudaStream_t stream);
cudaError_t mt_launch_build_rf_quads(const uint32_t* texels, int w, int h, int d, void* quads, cudaStream_t stream);
cudaError_t mt_launch_occupancy(const Tex3D& low, uint32_t* occ, float coverage, cudaStream_t stream);
cudaError_t mt_launch_reproject(const ReprojParams& P, cudaStream_t stream);
cudaError_t mt_launch_godrays(const GodRayParams& P, cudaStream_t stream);
cudaError_t mt_launch_mask_grey(const GodRayParams& P, float* out, cudaStream_t stream);
cudaError_t mt_launch_tonemap(const ToneMapParams& P, cudaStream_t stream);
cudaError_t mt_launch_txaa(const TxaaParams& P, cudaStream_t stream);
cudaError_t mt_launch_fma_probe(float* sink, int blocks, int iters, cudaStream_t stream);
