// fixed_kernels.h — argument blocks and host launchers of the shader-independent kernels (fixed_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include "cpvk_device.cuh"

#define CPVK_BIN_SMALL 16 /* primitives touching more tiles than this are binned by a whole CTA */

struct CpvkSetupArgs {
    const uint4* vsPos;
    cpvk_u32 nVerts, primCount;
    // vertex reuse (see CpvkDrawParams::vcache): raw vertex of stream position i = index[first + i] - lowest index
    const cpvk_u32* vcache;
    cpvk_u64 indexBuffer;
    cpvk_u32 indexStride, first;
    cpvk_u32 topology, frontFace, cullMode;
    float vpWidth, vpHeight;
    cpvk_i32 clipX0, clipY0, clipX1, clipY1;
    CpvkTriSetup* setups;
    CpvkBBox* bboxes;
    // binning pass 0 (count) is fused into setup: the bbox is in registers right here
    cpvk_u32 tilesX;
    cpvk_u32 tileRow0;   // first tile row of the render area (a band starts below row 0): tile ids count from there
    cpvk_u32* counts;    // [tiles], zeroed by the host
    cpvk_u32* largeList; // [primCount]
    cpvk_u32* meta;      // [2] = number of deferred (large) primitives, zeroed by the host
    // Single-pass binning (direct lists): every tile owns directCap slots at lists[tile * directCap]; the count pass claims
    // slots with the same atomic that counts and writes the primitive id there. A tile that needs more slots raises
    // meta[3] (counts stay exact), the rest of the draw is a no-op and the host replays count-exact binning from the scan on.
    cpvk_u32* directLists; // null = count only
    cpvk_u32 directCap;
    cpvk_u32 publish;      // 1 = the last CTA to finish stores the verdict to metaHost (no k_bin_large pass follows)
    cpvk_u32* ticket;      // CTA completion counter for `publish`, zeroed by the host
    cpvk_u32* metaHost;
};

struct CpvkBinArgs {
    const CpvkBBox* bboxes;
    cpvk_u32 primCount, tilesX, tilesY, tileRow0; // tilesY rows starting at tile row tileRow0
    cpvk_u32* counts;    // [tiles]
    cpvk_u32* offsets;   // [tiles + 1]
    cpvk_u32* cursors;   // [tiles]
    cpvk_u32* lists;     // [total entries]
    cpvk_u32* scratch;   // [total entries] or null (only needed by the long-list sort fallback)
    cpvk_u32* largeList; // [primCount]
    cpvk_u32* meta;      // [0] total entries, [1] longest list, [2] deferred-primitive count, [3] plan mismatch
    cpvk_u32* metaHost;  // the same four words in mapped host memory, stored by k_bin_scan
    // The launch plan the host committed to before knowing the counts; k_bin_scan checks it and raises meta[3].
    cpvk_u32 planCapacity;     // entries `lists` can hold
    cpvk_u32 planSortCap;      // longest list k_bin_sort was sized for (0 = not launched: lists must fit one raster chunk)
    cpvk_u32 planLargeCounted; // 1 = k_bin_large's count pass ran before the scan
    // single-pass binning of the deferred primitives (k_bin_large pass 2), see CpvkSetupArgs::directLists
    cpvk_u32* directLists;
    cpvk_u32 directCap;
    cpvk_u32* ticket;
};

struct CpvkClearArgs {
    float f[4];
    cpvk_u32 u[4];
    float depth;
    cpvk_u32 stencil;
    int isDepthStencil;
};

struct CpvkBlitArgs {
    CpvkDevAttachment src, dst;
    int srcX0, srcY0, srcX1, srcY1;
    int dstX0, dstY0, dstX1, dstY1;
    cpvk_u32 filter;
};

extern "C" {
cudaError_t cpvk_launch_setup(const CpvkSetupArgs* a, cudaStream_t s);
cudaError_t cpvk_launch_index_range(unsigned long long indexBuffer, unsigned indexStride, unsigned first, unsigned count, cpvk_u32* range /* [2] = {~lowest, highest}, zeroed by the caller */, cudaStream_t s);
cudaError_t cpvk_launch_bin(const CpvkBinArgs* a, int pass, int small, int large, cudaStream_t s); /* pass 0 = count, 1 = fill, 2 = count + write into the direct lists and publish the verdict (k_bin_large only); small/large select k_bin / k_bin_large (small primitives are counted by k_setup) */
cudaError_t cpvk_launch_bin_scan(const CpvkBinArgs* a, cudaStream_t s);
cudaError_t cpvk_launch_bin_sort(const CpvkBinArgs* a, unsigned capacity, cudaStream_t s);
cudaError_t cpvk_launch_clear(const CpvkDevAttachment* img, const CpvkClearArgs* c, cudaStream_t s);
cudaError_t cpvk_launch_copy_rows(unsigned long long dst, unsigned dstPitch, unsigned long long src, unsigned srcPitch, unsigned rowBytes, unsigned rows, cudaStream_t s);
cudaError_t cpvk_launch_blit(const CpvkBlitArgs* b, cudaStream_t s);
cudaError_t cpvk_launch_peer_barrier(const unsigned long long* flagArrays, unsigned count, unsigned self, unsigned sequence, const cpvk_u32* verdict /* null, or the binning words of an unsettled draw: [3] != 0 = do nothing, the host re-issues the barrier behind the replay */, cudaStream_t s);
cudaError_t cpvk_launch_selftest_div(const float* a, const float* b, unsigned n, float* shared, float* plain, cudaStream_t s);
}
