// stage_kernels.cu — the two sm_100a pipeline stages that call into application shaders.
// Compiled by nvcc to LTO-IR at build time; at vkCreateGraphicsPipelines time nvJitLink links this with the
// NVRTC-compiled shader translation unit (cpvk_vs_main / cpvk_fs_main / cpvk_spec_*), so the shader bodies and
// all baked pipeline state inline into these kernels — the GPU counterpart of the reference's per-pipeline
// JIT'd vertex and fragment wrappers (LLVMRuntime/PipelineCompiler.cpp:499-982, :991-1799).
//
//   cpvk_k_vertex : a1 + a3 of SURVEY §8(a) — input assembly + vertex fetch + VS + record store, one thread/index
//   cpvk_k_raster : a5 (pixel loop) a6 a7 a8 a9 a10 a11 a12 a13 — coverage, interpolation, FS, late depth/stencil,
//                   blend and format pack on a shared-memory tile; one CTA per 32x32 screen tile, API order kept.
#include "cpvk_device.cuh"

// ------------------------------------------------------------------------------------------------
// Vertex stage. ProcessInputAssembler[Indexed] (Draw.cpp:675-760): rawId = i,
// vertexId = firstVertex + i  or  vertexOffset + index[firstIndex + i]; the VS runs once per index (no cache).
extern "C" __global__ void __launch_bounds__(256) cpvk_k_vertex(const __grid_constant__ CpvkDrawParams p) {
    const cpvk_u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.count) return;
    cpvk_u32 vertexId;
    if (p.indexStride == 0) {
        vertexId = p.first + i;
    } else if (cpvk_vcache_on(p.vcache, p.count)) {
        // vertex reuse: thread i shades the i-th vertex of the draw's index range, whatever number of indices name it
        const cpvk_u32 lo = cpvk_vcache_lowest(p.vcache);
        if (i > p.vcache[1] - lo) return;
        vertexId = (cpvk_u32)p.vertexOffset + lo + i;
    } else {
        vertexId = (cpvk_u32)p.vertexOffset + cpvk_fetch_index(p.indexBuffer, p.indexStride, (cpvk_u64)p.first + i);
    }
    cpvk_vs_main(vertexId, p.instance, i, &p);
}

// ------------------------------------------------------------------------------------------------
// Fragment back end.

__device__ __forceinline__ bool cpvk_fcompare(float reference, float value, cpvk_u32 op) {
    // CompileFCompareTest (PipelineCompiler.cpp:1492-1508): ordered compares, NaN fails (NOT_EQUAL = ONE).
    switch (op) {
    case 0: return false;
    case 1: return reference < value;
    case 2: return reference == value;
    case 3: return reference <= value;
    case 4: return reference > value;
    case 5: return reference < value || reference > value;
    case 6: return reference >= value;
    default: return true;
    }
}
__device__ __forceinline__ bool cpvk_icompare(cpvk_u32 reference, cpvk_u32 value, cpvk_u32 op) {
    switch (op) {
    case 0: return false;
    case 1: return reference < value;
    case 2: return reference == value;
    case 3: return reference <= value;
    case 4: return reference > value;
    case 5: return reference != value;
    case 6: return reference >= value;
    default: return true;
    }
}
__device__ __forceinline__ cpvk_u32 cpvk_stencil_result(cpvk_u32 op, cpvk_u32 cur, cpvk_u32 ref) {
    // CompileGetStencilResult (PipelineCompiler.cpp:1382-1413); INC/DEC_CLAMP are *signed* i8 saturation there.
    switch (op) {
    case 0: return cur;
    case 1: return 0;
    case 2: return ref;
    case 3: { int v = (int)(signed char)cur + 1; if (v > 127) v = 127; return (cpvk_u32)v & 0xFFu; }
    case 4: { int v = (int)(signed char)cur - 1; if (v < -128) v = -128; return (cpvk_u32)v & 0xFFu; }
    case 5: return (~cur) & 0xFFu;
    case 6: return (cur + 1) & 0xFFu;
    default: return (cur - 1) & 0xFFu;
    }
}

// ApplyBlendFactor (Draw.cpp:956-1103)
__device__ __forceinline__ void cpvk_blend_factor(const float s[4], const float d[4], const float c[4], cpvk_u32 colourFactor,
                                                  cpvk_u32 alphaFactor, float v[4]) {
    v[0] = v[1] = v[2] = v[3] = 0.0f;
    switch (colourFactor) {
    case 0: break;
    case 1: v[0] = v[1] = v[2] = v[3] = 1.0f; break;
    case 2: for (int i = 0; i < 4; i++) v[i] = s[i]; break;
    case 3: for (int i = 0; i < 4; i++) v[i] = 1.0f - s[i]; break;
    case 4: for (int i = 0; i < 4; i++) v[i] = d[i]; break;
    case 5: for (int i = 0; i < 4; i++) v[i] = 1.0f - d[i]; break;
    case 6: v[0] = v[1] = v[2] = v[3] = s[3]; break;
    case 7: v[0] = v[1] = v[2] = v[3] = 1.0f - s[3]; break;
    case 8: v[0] = v[1] = v[2] = v[3] = d[3]; break;
    case 9: v[0] = v[1] = v[2] = v[3] = 1.0f - d[3]; break;
    case 10: for (int i = 0; i < 4; i++) v[i] = c[i]; break;
    case 11: for (int i = 0; i < 4; i++) v[i] = 1.0f - c[i]; break;
    case 12: v[0] = v[1] = v[2] = v[3] = c[3]; break;
    case 13: v[0] = v[1] = v[2] = v[3] = 1.0f - c[3]; break;
    case 14: { const float a = 1.0f - d[3]; const float f = a < s[3] ? a : s[3]; v[0] = v[1] = v[2] = f; v[3] = 1.0f; break; } // std::min(sa, 1-da)
    default: break;
    }
    if (colourFactor != alphaFactor) {
        switch (alphaFactor) {
        case 0: v[3] = 0.0f; break;
        case 1: case 14: v[3] = 1.0f; break;
        case 2: case 6: v[3] = s[3]; break;
        case 3: case 7: v[3] = 1.0f - s[3]; break;
        case 4: case 8: v[3] = d[3]; break;
        case 5: case 9: v[3] = 1.0f - d[3]; break;
        case 10: case 12: v[3] = c[3]; break;
        case 11: case 13: v[3] = 1.0f - c[3]; break;
        default: break;
        }
    }
}
// ApplyBlend (Draw.cpp:1105-1262)
__device__ __forceinline__ void cpvk_apply_blend(const float s[4], const float d[4], int a, float out[4]) {
    const int base = CPVK_SPEC_BLEND0 + a * 8;
    const cpvk_u32 srcC = cpvk_spec_u32(base + 1), dstC = cpvk_spec_u32(base + 2), opC = cpvk_spec_u32(base + 3);
    const cpvk_u32 srcA = cpvk_spec_u32(base + 4), dstA = cpvk_spec_u32(base + 5), opA = cpvk_spec_u32(base + 6);
    const float c[4] = {cpvk_spec_f32(0), cpvk_spec_f32(1), cpvk_spec_f32(2), cpvk_spec_f32(3)};
    float sf[4], df[4];
    cpvk_blend_factor(s, d, c, srcC, srcA, sf);
    cpvk_blend_factor(s, d, c, dstC, dstA, df);
    #pragma unroll
    for (int i = 0; i < 4; i++) {
        switch (opC) {
        case 0: out[i] = s[i] * sf[i] + d[i] * df[i]; break;
        case 1: out[i] = s[i] * sf[i] - d[i] * df[i]; break;
        case 2: out[i] = d[i] * df[i] - s[i] * sf[i]; break;
        case 3: out[i] = d[i] < s[i] ? d[i] : s[i]; break; // glm::min
        default: out[i] = s[i] < d[i] ? d[i] : s[i]; break; // glm::max
        }
    }
    if (opC != opA) {
        switch (opA) {
        case 0: out[3] = s[3] * sf[3] + d[3] * df[3]; break;
        case 1: out[3] = s[3] * sf[3] - d[3] * df[3]; break;
        case 2: out[3] = d[3] * df[3] - s[3] * sf[3]; break;
        case 3: out[3] = d[3] < s[3] ? d[3] : s[3]; break; // std::min
        default: out[3] = s[3] < d[3] ? d[3] : s[3]; break; // std::max
        }
    }
}

// (CPVK_RASTER_THREADS, not blockDim.x, strides the loops below: a compile-time stride spares a division for the trip count.)
// Copy `bytes` (a multiple of the texel size) between a global row segment and shared memory, all threads of
// the CTA cooperating over `rows` rows; 16-byte vectors when both sides allow, else 4-byte words, else bytes.
// `fullBytes` = bytes of a full-width tile row (a link-time constant per pipeline): when the row is full and everything is
// 16-byte aligned — every interior tile — the row/column split is a shift instead of a 32-bit division per vector.
__device__ __forceinline__ void cpvk_tile_copy(cpvk_u8* dst, cpvk_u32 dstPitch, const cpvk_u8* src, cpvk_u32 srcPitch, cpvk_u32 bytes, cpvk_u32 rows, cpvk_u32 fullBytes) {
    const cpvk_u64 align = ((cpvk_u64)dst | (cpvk_u64)src | dstPitch | srcPitch | bytes);
    if ((align & 15) == 0 && bytes == fullBytes && (fullBytes & (fullBytes - 1u)) == 0u) {
        const cpvk_u32 per = fullBytes >> 4; // power of two
        for (cpvk_u32 i = threadIdx.x; i < per * rows; i += CPVK_RASTER_THREADS) {
            const cpvk_u32 r = i / per, c = i & (per - 1u);
            reinterpret_cast<uint4*>(dst + (cpvk_u64)r * dstPitch)[c] = reinterpret_cast<const uint4*>(src + (cpvk_u64)r * srcPitch)[c];
        }
    } else if ((align & 15) == 0) {
        const cpvk_u32 per = bytes >> 4;
        for (cpvk_u32 i = threadIdx.x; i < per * rows; i += CPVK_RASTER_THREADS) {
            const cpvk_u32 r = i / per, c = i - r * per;
            reinterpret_cast<uint4*>(dst + (cpvk_u64)r * dstPitch)[c] = reinterpret_cast<const uint4*>(src + (cpvk_u64)r * srcPitch)[c];
        }
    } else if ((align & 3) == 0) {
        const cpvk_u32 per = bytes >> 2;
        for (cpvk_u32 i = threadIdx.x; i < per * rows; i += CPVK_RASTER_THREADS) {
            const cpvk_u32 r = i / per, c = i - r * per;
            reinterpret_cast<cpvk_u32*>(dst + (cpvk_u64)r * dstPitch)[c] = reinterpret_cast<const cpvk_u32*>(src + (cpvk_u64)r * srcPitch)[c];
        }
    } else {
        for (cpvk_u32 i = threadIdx.x; i < bytes * rows; i += CPVK_RASTER_THREADS) {
            const cpvk_u32 r = i / bytes, c = i - r * bytes;
            dst[(cpvk_u64)r * dstPitch + c] = src[(cpvk_u64)r * srcPitch + c];
        }
    }
}

// The same copy by ONE warp (lane-strided): a warp's region of a tile on its way back to HBM, without waiting for the other warps.
__device__ __forceinline__ void cpvk_region_copy(cpvk_u8* dst, cpvk_u32 dstPitch, const cpvk_u8* src, cpvk_u32 srcPitch, cpvk_u32 bytes, cpvk_u32 rows) {
    const cpvk_u32 lane = threadIdx.x & 31u;
    const cpvk_u64 align = ((cpvk_u64)dst | (cpvk_u64)src | dstPitch | srcPitch | bytes);
    if ((align & 15) == 0) {
        const cpvk_u32 per = bytes >> 4;
        for (cpvk_u32 i = lane; i < per * rows; i += 32u) {
            const cpvk_u32 r = i / per, c = i - r * per;
            reinterpret_cast<uint4*>(dst + (cpvk_u64)r * dstPitch)[c] = reinterpret_cast<const uint4*>(src + (cpvk_u64)r * srcPitch)[c];
        }
    } else if ((align & 3) == 0) {
        const cpvk_u32 per = bytes >> 2;
        for (cpvk_u32 i = lane; i < per * rows; i += 32u) {
            const cpvk_u32 r = i / per, c = i - r * per;
            reinterpret_cast<cpvk_u32*>(dst + (cpvk_u64)r * dstPitch)[c] = reinterpret_cast<const cpvk_u32*>(src + (cpvk_u64)r * srcPitch)[c];
        }
    } else {
        for (cpvk_u32 i = lane; i < bytes * rows; i += 32u) {
            const cpvk_u32 r = i / bytes, c = i - r * bytes;
            dst[(cpvk_u64)r * dstPitch + c] = src[(cpvk_u64)r * srcPitch + c];
        }
    }
}

// Fill a whole 32x32 shared-memory tile with one packed texel (texel size is a compile-time constant per pipeline).
CPVK_DEV void cpvk_tile_fill(cpvk_u8* dst, cpvk_u32 texel, const cpvk_u8* one) {
    if (texel == 4 || texel == 8 || texel == 2) { // 16-byte stores of the repeated texel: a 4 KB tile is one store per thread
        uint4 v;
        if (texel == 4) { const cpvk_u32 t = *reinterpret_cast<const cpvk_u32*>(one); v = make_uint4(t, t, t, t); }
        else if (texel == 8) { const uint2 t = *reinterpret_cast<const uint2*>(one); v = make_uint4(t.x, t.y, t.x, t.y); }
        else { const cpvk_u32 t = *reinterpret_cast<const unsigned short*>(one) * 0x10001u; v = make_uint4(t, t, t, t); }
        uint4* d = reinterpret_cast<uint4*>(dst);
        for (cpvk_u32 i = threadIdx.x; i < CPVK_TILE_W * CPVK_TILE_H * texel / 16; i += CPVK_RASTER_THREADS) d[i] = v;
        return;
    }
    for (cpvk_u32 i = threadIdx.x; i < CPVK_TILE_W * CPVK_TILE_H; i += CPVK_RASTER_THREADS) {
        cpvk_u8* d = dst + i * texel;
        if (texel == 16) *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(one);
        else if (texel == 8) *reinterpret_cast<uint2*>(d) = *reinterpret_cast<const uint2*>(one);
        else if (texel == 4) *reinterpret_cast<cpvk_u32*>(d) = *reinterpret_cast<const cpvk_u32*>(one);
        else if (texel == 2) *reinterpret_cast<unsigned short*>(d) = *reinterpret_cast<const unsigned short*>(one);
        else for (cpvk_u32 b = 0; b < texel; b++) d[b] = one[b];
    }
}

// x86 cvttss2si semantics for static_cast<int32_t>(float): NaN and out-of-range inputs give INT_MIN.
__device__ __forceinline__ int cpvk_cvtt(float v) { return (v >= 2147483648.0f || v < -2147483648.0f || v != v) ? (int)0x80000000 : (int)v; }

extern __shared__ __align__(16) cpvk_u8 cpvk_smem[];

// One CTA per screen tile. Warp w owns the 16x8 sub-rectangle (w&1, w>>1) of the tile, so no two warps ever touch
// the same pixel; inside a warp, triangles are taken strictly in list (= API) order and each triangle's candidate
// pixels are spread over the lanes, one pixel per lane per step, so a pixel sees its fragments in API order —
// the ordering the reference gets from its nested loops (Draw.cpp:1526-1593). Colour and depth live in shared
// memory in their *storage* format for the whole tile lifetime: every ROP is the reference's get/set-pixel
// round trip (GlslFunctions.cpp:842-928) on shared memory, and HBM sees one read and one write per tile byte.
#ifndef CPVK_COVER_ROWS
#define CPVK_COVER_ROWS 5   /* rows of the coverage window evaluated per pass over its columns (measured on C3/M1: 2 -> 242 us, 4 -> 235, 5 -> 228, 6 -> 231, 8 -> 249 with spills) */
#endif
#ifndef CPVK_ORDER_BITS
#define CPVK_ORDER_BITS (128 * CPVK_RASTER_THREADS) /* width of the id window the in-tile bitmap ordering covers: 128 bits per thread */
#endif
#ifndef CPVK_COVER_UNROLL
#define CPVK_COVER_UNROLL 1
#endif
#ifndef CPVK_RASTER_MIN_CTAS
#define CPVK_RASTER_MIN_CTAS 4 /* resident CTAs per SM the register allocation aims for; build.py builds 4, 3 and 2 */
#endif
extern "C" __global__ void __launch_bounds__(CPVK_RASTER_THREADS, CPVK_RASTER_MIN_CTAS) cpvk_k_raster(const __grid_constant__ CpvkDrawParams p) {
    const cpvk_u32 tx = blockIdx.x, tyr = blockIdx.y, tile = tyr * p.tilesX + tx; // grid = (tilesX, tilesY): no division to find the tile
    const bool triangles = cpvk_prim_vertices() == 3; // points and lines are not binned: every tile walks all of them (below)
    if (triangles && p.binMeta[3] != 0) return; // the speculative launch plan did not fit this draw: the host replays it
    const bool listsSorted = triangles && p.binMeta[1] > CPVK_ORDER_MAX;
    cpvk_u32 listBegin = 0u, listEnd = 1u;
    if (triangles) {
        if (p.directCap) { listBegin = tile * p.directCap; listEnd = listBegin + min(p.tileOffsets[tile], p.directCap); } // grids of up to 2^32 / directCap tiles (cpvk_abi.cpp keeps to that)
        else { listBegin = p.tileOffsets[tile]; listEnd = p.tileOffsets[tile + 1]; }
    }
    const cpvk_u32 lazyMask = p.lazyMask;
    const int tileX0 = (int)tx * CPVK_TILE_W, tileY0 = (int)(tyr + p.tileRow0) * CPVK_TILE_H;
    // nothing to draw and no clear to fold: done — unless mirrors are on, then even untouched tiles of the band travel
    if (listBegin == listEnd && lazyMask == 0 && p.mirrorCount == 0) return;

    const cpvk_u32 dsFormat = cpvk_spec_u32(CPVK_SPEC_DS_FORMAT);
    const bool depthTest = cpvk_spec_u32(CPVK_SPEC_DEPTH_TEST) != 0, depthWrite = cpvk_spec_u32(CPVK_SPEC_DEPTH_WRITE) != 0;
    const bool boundsTest = cpvk_spec_u32(CPVK_SPEC_BOUNDS_TEST) != 0, stencilTest = cpvk_spec_u32(CPVK_SPEC_STENCIL_TEST) != 0;
    const cpvk_u32 depthOp = cpvk_spec_u32(CPVK_SPEC_DEPTH_OP);
    const int colorCount = (int)cpvk_spec_u32(CPVK_SPEC_COLOR_COUNT);
    const bool fmtDepth = dsFormat != 0 && dsFormat != 127, fmtStencil = dsFormat >= 127 && dsFormat <= 130 && dsFormat != 0;
    const bool stencilOn = stencilTest && fmtStencil;
    const bool dsUsed = dsFormat != 0 && p.ds.address != 0 && (depthTest || boundsTest || stencilOn);

    // ---- shared-memory layout: [xf 32][yf 32] | depth tile | colour tiles ----
    float* sXf = reinterpret_cast<float*>(cpvk_smem);
    float* sYf = sXf + CPVK_TILE_W;
    float* sLut = sYf + CPVK_TILE_H; // [256] (float)k / 255.0f, see cpvk_get_pixel_f32_dyn
    cpvk_u32 smemOff = (CPVK_TILE_W + CPVK_TILE_H + 256) * 4;
    const cpvk_u32 dsTexel = dsFormat ? cpvk_texel_size(dsFormat) : 0;
    cpvk_u8* sDepth = cpvk_smem + smemOff;
    const cpvk_u32 dsPitch = dsTexel * CPVK_TILE_W;
    if (dsUsed) smemOff += ((dsPitch * CPVK_TILE_H) + 15u) & ~15u;
    cpvk_u8* sColor[CPVK_MAX_COLOR];
    cpvk_u32 cTexel[CPVK_MAX_COLOR];
    #pragma unroll
    for (int a = 0; a < CPVK_MAX_COLOR; a++) {
        sColor[a] = nullptr; cTexel[a] = 0;
        if (a < colorCount) {
            const cpvk_u32 cf = cpvk_spec_u32(CPVK_SPEC_COLOR_FORMAT0 + a);
            if (cf != 0 && p.color[a].address != 0) {
                cTexel[a] = cpvk_texel_size(cf);
                sColor[a] = cpvk_smem + smemOff;
                smemOff += (cTexel[a] * CPVK_TILE_W * CPVK_TILE_H + 15u) & ~15u;
            }
        }
    }

    // tile extent inside the render area
    const int x1 = min(tileX0 + CPVK_TILE_W, p.clipX1), y1 = min(tileY0 + CPVK_TILE_H, p.clipY1);
    // rows of this tile inside the render area: a band need not start on a tile row, and the rows above it belong to
    // another GPU (which stores them into this frame while this kernel runs) — they are neither read nor written here
    const int wy0 = max(tileY0, p.clipY0);
    const int tw = x1 - tileX0, th = y1 - wy0;
    if (tw <= 0 || th <= 0) return;
    const cpvk_u32 skipRows = (cpvk_u32)(wy0 - tileY0);

    // pixel centres in NDC, exactly as Draw.cpp:1524,1573,1578: ((float)x / W + (1/W)*0.5) * 2 - 1
    if (threadIdx.x < CPVK_TILE_W) {
        const float W = p.vpWidth; const float hp = (1.0f / W) * 0.5f;
        sXf[threadIdx.x] = ((float)(tileX0 + (int)threadIdx.x) / W + hp) * 2.0f - 1.0f;
    } else if (threadIdx.x < CPVK_TILE_W + CPVK_TILE_H) {
        const int j = (int)threadIdx.x - CPVK_TILE_W;
        const float H = p.vpHeight; const float hp = (1.0f / H) * 0.5f;
        sYf[j] = ((float)(tileY0 + j) / H + hp) * 2.0f - 1.0f;
    }
    sLut[threadIdx.x] = cpvk_unorm8(threadIdx.x); // == (float)k / 255.0f for every k; CPVK_RASTER_THREADS == 256
    // Unsorted lists hold at most CPVK_ORDER_MAX ids (the host sorts otherwise), two per thread: fetch them now so the load
    // overlaps tile staging, and park them where the ordering pass below expects them — the staging barrier then covers both.
    cpvk_u32 keyA = 0xFFFFFFFFu, keyB = 0xFFFFFFFFu;
    if (triangles && !listsSorted) {
        const cpvk_u32 nList = listEnd - listBegin;
        if (threadIdx.x < nList) keyA = __ldg(p.tileLists + listBegin + threadIdx.x);
        if (threadIdx.x + CPVK_CHUNK < nList) keyB = __ldg(p.tileLists + listBegin + CPVK_CHUNK + threadIdx.x);
    }
    // ---- stage the tile: HBM -> shared, or the packed clear value when a deferred clear is folded into this draw ----
    if (dsUsed) {
        if (lazyMask & 0x100u) {
            __align__(16) cpvk_u8 one[16];
            cpvk_set_depth_stencil(dsFormat, one, p.lazyDepth, p.lazyStencil);
            cpvk_tile_fill(sDepth, dsTexel, one);
        } else
            cpvk_tile_copy(sDepth + skipRows * dsPitch, dsPitch, reinterpret_cast<const cpvk_u8*>(p.ds.address) + (cpvk_u64)wy0 * p.ds.rowPitch + (cpvk_u64)tileX0 * dsTexel,
                           p.ds.rowPitch, (cpvk_u32)tw * dsTexel, (cpvk_u32)th, dsPitch);
    }
    #pragma unroll
    for (int a = 0; a < CPVK_MAX_COLOR; a++)
        if (sColor[a]) {
            if (lazyMask & (1u << a)) {
                const cpvk_u32 cf = cpvk_spec_u32(CPVK_SPEC_COLOR_FORMAT0 + a);
                __align__(16) cpvk_u8 one[16];
                if (cpvk_format_is_int(cf)) { const cpvk_u32 v[4] = {p.lazyColor[a][0], p.lazyColor[a][1], p.lazyColor[a][2], p.lazyColor[a][3]}; cpvk_set_pixel_int(cf, one, v); }
                else { const float v[4] = {__uint_as_float(p.lazyColor[a][0]), __uint_as_float(p.lazyColor[a][1]), __uint_as_float(p.lazyColor[a][2]), __uint_as_float(p.lazyColor[a][3])}; cpvk_set_pixel_f32(cf, one, v); }
                cpvk_tile_fill(sColor[a], cTexel[a], one);
            } else
                cpvk_tile_copy(sColor[a] + skipRows * cTexel[a] * CPVK_TILE_W, cTexel[a] * CPVK_TILE_W, reinterpret_cast<const cpvk_u8*>(p.color[a].address) + (cpvk_u64)wy0 * p.color[a].rowPitch + (cpvk_u64)tileX0 * cTexel[a],
                               p.color[a].rowPitch, (cpvk_u32)tw * cTexel[a], (cpvk_u32)th, cTexel[a] * CPVK_TILE_W);
        }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // warp region (CPVK_REGION_W wide, CPVK_REGION_H tall) clipped to the tile extent
    const int rx0 = tileX0 + (warp & 1) * CPVK_REGION_W, ry0 = tileY0 + (warp >> 1) * CPVK_REGION_H;
    const int rx1 = min(rx0 + CPVK_REGION_W, x1), ry1 = min(ry0 + CPVK_REGION_H, y1);
    cpvk_u32 nCov = 0, nPass = 0;

    // Shared-memory scratch after the tiles: the staged triangle chunk (six uint4 planes of the setup records + bboxes,
    // read by every warp of the tile) and one compacted hit list per warp.
    uint4* sQ = reinterpret_cast<uint4*>(cpvk_smem + smemOff);                                   // [6][CPVK_CHUNK]
    uint2* sBB = reinterpret_cast<uint2*>(cpvk_smem + smemOff + 6 * CPVK_CHUNK * 16);            // [CPVK_CHUNK]
    cpvk_u8* sHit = cpvk_smem + smemOff + 6 * CPVK_CHUNK * 16 + CPVK_CHUNK * 8 + warp * CPVK_CHUNK; // [warps][CPVK_CHUNK] chunk-local ids
    unsigned short* sFrag = reinterpret_cast<unsigned short*>(cpvk_smem + smemOff + 6 * CPVK_CHUNK * 16 + CPVK_CHUNK * 8 + (CPVK_RASTER_THREADS / 32) * CPVK_CHUNK)
                            + warp * CPVK_FRAG_CAP;                                                  // [warps][CPVK_FRAG_CAP]
    cpvk_u8* sMask = cpvk_smem + smemOff + 6 * CPVK_CHUNK * 16 + CPVK_CHUNK * 8 + (CPVK_RASTER_THREADS / 32) * (CPVK_CHUNK + CPVK_FRAG_CAP * 2); // [CPVK_CHUNK] warp regions a bbox meets
    cpvk_u32* sSorted = reinterpret_cast<cpvk_u32*>(sMask + CPVK_CHUNK); // [CPVK_ORDER_MAX] the tile's whole list in API order, read chunk by chunk
    // Ordering scratch (aliases the setup planes, idle until the chunk is staged): a bitmap of the ids present in this tile's
    // list relative to the lowest one, the warps' lowest / highest ids, the warps' bit counts.
    cpvk_u32* sBits = reinterpret_cast<cpvk_u32*>(sQ);   // [CPVK_ORDER_BITS / 32]
    cpvk_u32* sRange = sBits + CPVK_ORDER_BITS / 32;     // [0..8) lowest id per warp, [8..16) highest, [16..24) set bits per warp
    if (triangles && !listsSorted) {
        reinterpret_cast<cpvk_u32*>(sBB)[threadIdx.x] = keyA; reinterpret_cast<cpvk_u32*>(sBB)[CPVK_CHUNK + threadIdx.x] = keyB; // sKeys of the ranking pass (fallback): CPVK_ORDER_MAX words = the bbox plane
        reinterpret_cast<uint4*>(sBits)[threadIdx.x] = make_uint4(0u, 0u, 0u, 0u); // CPVK_ORDER_BITS == 128 bits per thread
        const cpvk_u32 mn = __reduce_min_sync(0xFFFFFFFFu, min(keyA, keyB));
        const cpvk_u32 mx = __reduce_max_sync(0xFFFFFFFFu, max(keyA == 0xFFFFFFFFu ? 0u : keyA, keyB == 0xFFFFFFFFu ? 0u : keyB));
        if (lane == 0) { sRange[warp] = mn; sRange[8 + warp] = mx; }
    }
    __syncthreads(); // tile, pixel centres, lut and the unsorted ids are staged

    // ---- the fragment wrapper epilogue (PipelineCompiler.cpp:1061-1080) on the shared tile; returns "colour written" ----
    auto rop = [&](int px, int py, float fragDepth, bool front, const CpvkFragOut& out) -> bool {
        cpvk_u8* dsp = sDepth + (cpvk_u32)py * dsPitch + (cpvk_u32)px * dsTexel;
        float currentDepth = 0.0f; cpvk_u32 currentStencil = 0;
        if ((boundsTest || depthTest) && fmtDepth && dsUsed) currentDepth = cpvk_get_depth(dsFormat, dsp);
        if (stencilOn && dsUsed) currentStencil = cpvk_get_stencil(dsFormat, dsp);
        if (boundsTest && dsFormat != 0) // FCmpULT / FCmpUGT: unordered counts as out of bounds
            if (!((currentDepth >= cpvk_spec_f32(4)) && (currentDepth <= cpvk_spec_f32(5)))) return false;
        bool stencilResult = true, depthResult = true;
        cpvk_u32 sref = 0;
        if (stencilOn) {
            const int sb = front ? CPVK_SPEC_STENCIL_FRONT : CPVK_SPEC_STENCIL_BACK;
            sref = cpvk_spec_u32(sb + 6) & 0xFFu;
            const cpvk_u32 cmask = cpvk_spec_u32(sb + 4) & 0xFFu;
            stencilResult = cpvk_icompare(sref & cmask, currentStencil & cmask, cpvk_spec_u32(sb + 3));
        }
        if (depthTest && dsFormat != 0 && dsFormat != 127) depthResult = cpvk_fcompare(fragDepth, currentDepth, depthOp);
        if (stencilOn) {
            // write ops always come from the FRONT state: both arms call depthFunctions(true)
            // (PipelineCompiler.cpp:1175-1185, reference defect kept for parity)
            const int wb = CPVK_SPEC_STENCIL_FRONT;
            const cpvk_u32 failR = cpvk_stencil_result(cpvk_spec_u32(wb + 0), currentStencil, sref);
            const cpvk_u32 passR = cpvk_stencil_result(cpvk_spec_u32(wb + 1), currentStencil, sref);
            const cpvk_u32 dfailR = cpvk_stencil_result(cpvk_spec_u32(wb + 2), currentStencil, sref);
            cpvk_u32 wv = stencilResult ? (depthResult ? passR : dfailR) : failR;
            const cpvk_u32 wmask = cpvk_spec_u32(wb + 5) & 0xFFu;
            wv = (wv & wmask) | (currentStencil & (~wmask & 0xFFu));
            if (dsUsed) {
                if (depthResult && depthTest && depthWrite && dsFormat != 127) cpvk_set_depth_stencil(dsFormat, dsp, fragDepth, wv);
                else cpvk_set_depth_stencil(dsFormat, dsp, fmtDepth ? cpvk_get_depth(dsFormat, dsp) : 0.0f, wv); // GlslFunctions.cpp:898-914
            }
        } else if (depthTest && depthWrite && dsFormat != 0 && dsFormat != 127) {
            if (depthResult && dsUsed) // SetDepthPixelXXX keeps the stencil byte (GlslFunctions.cpp:880-896)
                cpvk_set_depth_stencil(dsFormat, dsp, fragDepth, fmtStencil ? cpvk_get_stencil(dsFormat, dsp) : 0u);
        }
        if (!(stencilResult && depthResult)) return false;
        #pragma unroll
        for (int a = 0; a < CPVK_MAX_COLOR; a++) {
            if (!sColor[a]) continue;
            const cpvk_u32 cf = cpvk_spec_u32(CPVK_SPEC_COLOR_FORMAT0 + a);
            cpvk_u8* cp = sColor[a] + ((cpvk_u32)py * CPVK_TILE_W + (cpvk_u32)px) * cTexel[a];
            const int bb8 = CPVK_SPEC_BLEND0 + a * 8;
            const bool blendOn = cpvk_spec_u32(bb8) != 0;
            const cpvk_u32 wm = cpvk_spec_u32(bb8 + 7);
            if (cpvk_format_is_int(cf)) {
                cpvk_u32 v[4] = {out.color[a][0], out.color[a][1], out.color[a][2], out.color[a][3]};
                if (wm != 0xFu) { cpvk_u32 d[4]; cpvk_get_pixel_int(cf, cp, d); for (int k = 0; k < 4; k++) if (!(wm & (1u << k))) v[k] = d[k]; }
                cpvk_set_pixel_int(cf, cp, v);
            } else {
                float v[4] = {__uint_as_float(out.color[a][0]), __uint_as_float(out.color[a][1]), __uint_as_float(out.color[a][2]), __uint_as_float(out.color[a][3])};
                if (blendOn || wm != 0xFu) {
                    float d[4]; // ImageFetch of the destination (Draw.cpp:1283-1298)
                    if (cf == 37 || cf == 44) cpvk_get_pixel_f32_dyn(cf, cp, d, sLut); else cpvk_get_pixel_f32(cf, cp, d);
                    if (blendOn) { float r[4]; cpvk_apply_blend(v, d, a, r); v[0] = r[0]; v[1] = r[1]; v[2] = r[2]; v[3] = r[3]; }
                    for (int k = 0; k < 4; k++) if (!(wm & (1u << k))) v[k] = d[k]; // PipelineCompiler.cpp:1695-1698
                }
                cpvk_set_pixel_f32(cf, cp, v);
            }
        }
        return true;
    };


    // ---- shade + ROP one batch: lane = one fragment (active, chunk-local triangle kt, tile-local pixel px,py) ----
    // Weights are recomputed from the staged edges (EdgeFunction, Draw.cpp:415-418 — same expression as the coverage
    // test, so the same bits), then GetFragmentInput's normalisation/depth, DrawPixel's viewport depth, the fragment
    // shader, and the epilogue. Fragments of one batch are in API order by lane; the ROP of fragments that share a
    // pixel is serialised lowest lane first (a pixel belongs to exactly one warp, so that is the only hazard).
    auto shadeBatch = [&](bool active, cpvk_u32 kt, int px, int py, bool distinct) {
        CpvkFragOut out;
        bool survive = false, front = true;
        float fragDepth = 0.0f;
        cpvk_u32 key = 0x80000000u | (cpvk_u32)lane; // unique for idle lanes
        float w0 = 0.0f, w1 = 0.0f, w2 = 0.0f;
        if (active) {
            // EdgeFunction at the pixel centre; inside = none of the three is < 0 (no fill rule; NaN accepted). Packed
            // fragments were already found covered by the same expression, so for them this is a no-op.
            const uint4 q0 = sQ[kt], q1 = sQ[CPVK_CHUNK + kt], q2 = sQ[2 * CPVK_CHUNK + kt];
            const float xf = sXf[px], yf = sYf[py];
            w0 = (xf - __uint_as_float(q0.x)) * __uint_as_float(q0.z) - (yf - __uint_as_float(q0.y)) * __uint_as_float(q0.w);
            w1 = (xf - __uint_as_float(q1.x)) * __uint_as_float(q1.z) - (yf - __uint_as_float(q1.y)) * __uint_as_float(q1.w);
            w2 = (xf - __uint_as_float(q2.x)) * __uint_as_float(q2.z) - (yf - __uint_as_float(q2.y)) * __uint_as_float(q2.w);
            active = !(w0 < 0.0f || w1 < 0.0f || w2 < 0.0f);
        }
        const cpvk_u32 covMask = __ballot_sync(0xFFFFFFFFu, active);
        if (covMask == 0) return;
        nCov += __popc(covMask);
        if (active) {
            const uint4 q3 = sQ[3 * CPVK_CHUNK + kt], q4 = sQ[4 * CPVK_CHUNK + kt], q5 = sQ[5 * CPVK_CHUNK + kt];
            const float area = __uint_as_float(q3.w);
            CpvkFragCtx ctx;
            { float w[3] = {w0, w1, w2}; cpvk_div_shared(w, area); w0 = w[0]; w1 = w[1]; w2 = w[2]; } // w /= area (Draw.cpp:905-907), one reciprocal for the three
            const float depth = __uint_as_float(q3.x) * w0 + __uint_as_float(q3.y) * w1 + __uint_as_float(q3.z) * w2; // Draw.cpp:909
            ctx.w[0] = w0; ctx.w[1] = w1; ctx.w[2] = w2;
            ctx.pw[0] = __uint_as_float(q4.x); ctx.pw[1] = __uint_as_float(q4.y); ctx.pw[2] = __uint_as_float(q4.z);
            ctx.unitW = ctx.pw[0] == 1.0f && ctx.pw[1] == 1.0f && ctx.pw[2] == 1.0f;
            {
                float den = 0.0f; // dead code unless the shader has a perspective-interpolated input
                if (ctx.unitW) { den += w0; den += w1; den += w2; }
                else { den += w0 / ctx.pw[0]; den += w1 / ctx.pw[1]; den += w2 / ctx.pw[2]; } // (ptxas merges these reciprocals with the interpolation's when it can)
                ctx.persDen = den;
            }
            front = (q4.w & 1u) != 0;
            ctx.idx[0] = q5.x; ctx.idx[1] = q5.y; ctx.idx[2] = q5.z; ctx.provoking = q5.w;
            const int x = tileX0 + px, y = tileY0 + py;
            ctx.fragCoord[0] = cpvk_spec_u32(CPVK_SPEC_ORIGIN_UPPER) ? (float)x : p.vpWidth - (float)x - 1.0f; // Draw.cpp:1579
            ctx.fragCoord[1] = (float)y; ctx.fragCoord[2] = depth; ctx.fragCoord[3] = 1.0f;
            ctx.v[0] = p.vsOut + (cpvk_u64)q5.x * p.vsStride; ctx.v[1] = p.vsOut + (cpvk_u64)q5.y * p.vsStride;
            ctx.v[2] = p.vsOut + (cpvk_u64)q5.z * p.vsStride; ctx.vProv = p.vsOut + (cpvk_u64)q5.w * p.vsStride;
            ctx.dp = &p; ctx.unorm8 = sLut;
            fragDepth = (p.vpMaxDepth - p.vpMinDepth) * depth + p.vpMinDepth;                       // DrawPixel, Draw.cpp:1310
            survive = !cpvk_fs_main(&ctx, &out);
            key = (cpvk_u32)(py * CPVK_TILE_W + px);
        }
        // a large triangle's batch holds one pixel per lane by construction (warp-uniform flag): nothing to serialise
        const cpvk_u32 same = distinct ? (1u << lane) : __match_any_sync(0xFFFFFFFFu, key);
        cpvk_u32 pending = __ballot_sync(0xFFFFFFFFu, survive);
        const cpvk_u32 lowerMask = (1u << lane) - 1u;
        cpvk_u32 writtenMask = 0;
        #pragma unroll 1
        while (pending) {
            const bool go = ((pending >> lane) & 1u) && ((same & pending & lowerMask) == 0u);
            bool wrote = false;
            if (go) wrote = rop(px, py, fragDepth, front, out);
            pending &= ~__ballot_sync(0xFFFFFFFFu, go);
            writtenMask |= __ballot_sync(0xFFFFFFFFu, wrote);
            __syncwarp();
        }
        nPass += __popc(writtenMask);
    };

    const bool regionLive = rx0 < rx1 && ry0 < ry1;
    if (!triangles) {
        // ---- points and lines (ProcessPoints / ProcessLines, Draw.cpp:1315-1508) ----
        // The reference tests every pixel of the viewport against every line; here each tile walks all primitives in API
        // order, 256 at a time: one thread derives one primitive's terms into shared memory, then every lane tests its own
        // four pixels of the warp's region against primitives whose (conservative) pixel box touches the region. A pixel
        // belongs to one lane for the whole draw, so its fragments reach the ROP in API order by construction.
        const bool points = cpvk_spec_u32(CPVK_SPEC_TOPOLOGY) == 0u;
        float* sRec = reinterpret_cast<float*>(sQ); // [CPVK_CHUNK][24]
        const bool useCache = p.indexStride != 0 && cpvk_vcache_on(p.vcache, p.count);
        auto slotOf = [&](cpvk_u32 i) -> cpvk_u32 { return useCache ? cpvk_fetch_index(p.indexBuffer, p.indexStride, (cpvk_u64)p.first + i) - cpvk_vcache_lowest(p.vcache) : i; };
        #pragma unroll 1
        for (cpvk_u32 base = 0; base < p.primCount; base += CPVK_CHUNK) {
            const int n = (int)min((cpvk_u32)CPVK_CHUNK, p.primCount - base);
            __syncthreads();
            if ((int)threadIdx.x < n) {
                const cpvk_u32 prim = base + threadIdx.x;
                float* r = sRec + threadIdx.x * 24;
                int bx0, by0, bx1, by1;
                const float W = p.vpWidth, H = p.vpHeight;
                if (points) {
                    const cpvk_u32 s0 = slotOf(prim);
                    const uint4 pv = __ldg(p.vsPos + s0);
                    const float pw = __uint_as_float(pv.w);
                    const float X = __uint_as_float(pv.x), Y = __uint_as_float(pv.y), Z = __uint_as_float(pv.z); // stored divided by w (cpvk_store_position)
                    const float pointSize = __uint_as_float(__ldg(p.vsPointSize + s0));
                    const int sx = cpvk_cvtt((X + 1.0f) * 0.5f * (W - 1.0f)), sy = cpvk_cvtt((Y + 1.0f) * 0.5f * (H - 1.0f));
                    const int half = cpvk_cvtt(ceilf(pointSize / 2.0f));
                    bx0 = max(0, sx - half); by0 = max(0, sy - half);                               // Draw.cpp:1348-1351: the pixel loop's exact domain
                    bx1 = min(cpvk_cvtt(W), sx + half + 1); by1 = min(cpvk_cvtt(H), sy + half + 1);
                    r[0] = __int_as_float(sx); r[1] = __int_as_float(sy); r[2] = pointSize; r[3] = Z; r[4] = pw; r[17] = __uint_as_float(s0); r[18] = __uint_as_float(s0);
                } else {
                    const cpvk_u32 i0 = cpvk_spec_u32(CPVK_SPEC_TOPOLOGY) == 1u ? prim * 2u : prim;
                    const cpvk_u32 s0 = slotOf(i0), s1 = slotOf(i0 + 1u);
                    const uint4 v0 = __ldg(p.vsPos + s0), v1 = __ldg(p.vsPos + s1);
                    const float w0 = __uint_as_float(v0.w), w1 = __uint_as_float(v1.w);
                    const float P0[4] = {__uint_as_float(v0.x), __uint_as_float(v0.y), __uint_as_float(v0.z), w0}; // stored divided by w (cpvk_store_position)
                    const float P1[4] = {__uint_as_float(v1.x), __uint_as_float(v1.y), __uint_as_float(v1.z), w1};
                    const float lw = cpvk_spec_f32(6);
                    const float lwx = lw / W, lwy = lw / H;
                    const float d0 = P1[0] - P0[0], d1 = P1[1] - P0[1], d2 = P1[2] - P0[2], d3 = P1[3] - P0[3];
                    const float sqr = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;  // glm::normalize(vec4): x * (1 / sqrt(dot))
                    const float inv = 1.0f / sqrtf(sqr);
                    const float dirx = d0 * inv, diry = d1 * inv;
                    const float ox = diry * lwx, oy = (-dirx) * lwy;           // perpendicular (dir.y, -dir.x) * lineWidth, component-wise
                    r[0] = P0[0] + ox; r[1] = P0[1] + oy; r[2] = P0[0] - ox; r[3] = P0[1] - oy; // p00, p01
                    r[4] = P1[0] + ox; r[5] = P1[1] + oy; r[6] = P1[0] - ox; r[7] = P1[1] - oy; // p10, p11
                    r[8] = P0[0]; r[9] = P0[1]; r[10] = d0; r[11] = d1;
                    const float len = sqrtf(d0 * d0 + d1 * d1);                 // glm::length(vec2)
                    r[12] = len * len; r[13] = P0[2]; r[14] = P1[2]; r[15] = w0; r[16] = w1;
                    r[17] = __uint_as_float(s0); r[18] = __uint_as_float(s1);
                    // Pixel box: exact for a proper parallelogram (+2 px for rounding); a degenerate or non-finite quad can pass
                    // the four >= 0 tests anywhere (all-zero edge functions), so it keeps the whole viewport like the reference.
                    bx0 = 0; by0 = 0; bx1 = cpvk_cvtt(ceilf(W)); by1 = cpvk_cvtt(ceilf(H));
                    const float cross = (r[2] - r[0]) * (r[5] - r[1]) - (r[3] - r[1]) * (r[4] - r[0]);
                    float mnx = fminf(fminf(r[0], r[2]), fminf(r[4], r[6])), mxx = fmaxf(fmaxf(r[0], r[2]), fmaxf(r[4], r[6]));
                    float mny = fminf(fminf(r[1], r[3]), fminf(r[5], r[7])), mxy = fmaxf(fmaxf(r[1], r[3]), fmaxf(r[5], r[7]));
                    const float sum = ((r[0] + r[2]) + (r[4] + r[6])) + ((r[1] + r[3]) + (r[5] + r[7]));
                    if (fabsf(cross) > 0.0f && fabsf(sum) < 1e30f && fabsf(cross) < 1e30f) { // finite corners, non-zero area
                        const float fx0 = (mnx + 1.0f) * 0.5f * W - 2.0f, fx1 = (mxx + 1.0f) * 0.5f * W + 3.0f;
                        const float fy0 = (mny + 1.0f) * 0.5f * H - 2.0f, fy1 = (mxy + 1.0f) * 0.5f * H + 3.0f;
                        bx0 = max(bx0, cpvk_cvtt(floorf(fmaxf(fx0, -1.0f)))); bx1 = min(bx1, cpvk_cvtt(ceilf(fminf(fx1, 40000.0f))));
                        by0 = max(by0, cpvk_cvtt(floorf(fmaxf(fy0, -1.0f)))); by1 = min(by1, cpvk_cvtt(ceilf(fminf(fy1, 40000.0f))));
                    }
                }
                bx0 = max(bx0, p.clipX0); by0 = max(by0, p.clipY0); bx1 = min(bx1, p.clipX1); by1 = min(by1, p.clipY1);
                r[19] = __int_as_float(bx0); r[20] = __int_as_float(by0); r[21] = __int_as_float(bx1); r[22] = __int_as_float(by1);
            }
            __syncthreads();
            if (!regionLive) continue;
            #pragma unroll 1
            for (int j = 0; j < n; j++) {
                const float* r = sRec + j * 24;
                const int bx0 = __float_as_int(r[19]), by0 = __float_as_int(r[20]), bx1 = __float_as_int(r[21]), by1 = __float_as_int(r[22]);
                if (!(bx0 < rx1 && bx1 > rx0 && by0 < ry1 && by1 > ry0)) continue; // warp-uniform
                #pragma unroll 1
                for (int k = 0; k < CPVK_REGION_W * CPVK_REGION_H / 32; k++) { // 32 pixels of the region per step: whole rows of it
                    const int x = rx0 + (lane & (CPVK_REGION_W - 1)), y = ry0 + lane / CPVK_REGION_W + (32 / CPVK_REGION_W) * k;
                    bool covered = x >= bx0 && x < bx1 && y >= by0 && y < by1 && x < rx1 && y < ry1;
                    float t = 0.0f;
                    if (covered) {
                        if (points) {
                            const float ps = r[2];
                            const float sx = 0.5f + (float)(x - __float_as_int(r[0])) / ps, sy = 0.5f + (float)(y - __float_as_int(r[1])) / ps;
                            covered = sx >= 0.0f && sy >= 0.0f && sx <= 1.0f && sy <= 1.0f;
                        } else {
                            const float xf = sXf[x - tileX0], yf = sYf[y - tileY0];
                            // EdgeFunction(a, b, c) = (c.x - a.x) * (b.y - a.y) - (c.y - a.y) * (b.x - a.x), Draw.cpp:410-413
                            const float e0 = (xf - r[0]) * (r[3] - r[1]) - (yf - r[1]) * (r[2] - r[0]); // (p00, p01)
                            const float e1 = (xf - r[6]) * (r[5] - r[7]) - (yf - r[7]) * (r[4] - r[6]); // (p11, p10)
                            const float e2 = (xf - r[4]) * (r[1] - r[5]) - (yf - r[5]) * (r[0] - r[4]); // (p10, p00)
                            const float e3 = (xf - r[2]) * (r[7] - r[3]) - (yf - r[3]) * (r[6] - r[2]); // (p01, p11)
                            covered = e0 >= 0.0f && e1 >= 0.0f && e2 >= 0.0f && e3 >= 0.0f;
                            t = ((xf - r[8]) * r[10] + (yf - r[9]) * r[11]) / r[12];
                        }
                    }
                    const cpvk_u32 covMask = __ballot_sync(0xFFFFFFFFu, covered);
                    if (covMask == 0) continue;
                    nCov += __popc(covMask);
                    CpvkFragOut out;
                    bool wrote = false;
                    if (covered) {
                        CpvkFragCtx ctx;
                        float depth;
                        if (points) {
                            ctx.w[0] = 1.0f; ctx.w[1] = 0.0f; ctx.w[2] = 0.0f; ctx.pw[0] = r[4]; ctx.pw[1] = 1.0f; ctx.pw[2] = 1.0f;
                            depth = r[3];
                        } else {
                            ctx.w[0] = 1.0f - t; ctx.w[1] = t; ctx.w[2] = 0.0f; ctx.pw[0] = r[15]; ctx.pw[1] = r[16]; ctx.pw[2] = 1.0f;
                            depth = r[13] * t + r[14] * (1.0f - t); // Draw.cpp:1494: the depth weights are the attribute weights swapped
                        }
                        ctx.unitW = ctx.pw[0] == 1.0f && ctx.pw[1] == 1.0f;
                        { float den = 0.0f;
                          if (ctx.unitW) { den += ctx.w[0]; den += ctx.w[1]; } else { den += ctx.w[0] / ctx.pw[0]; den += ctx.w[1] / ctx.pw[1]; }
                          ctx.persDen = den; }
                        const cpvk_u32 s0 = __float_as_uint(r[17]), s1 = __float_as_uint(r[18]);
                        ctx.idx[0] = s0; ctx.idx[1] = s1; ctx.idx[2] = s1; ctx.provoking = s0;
                        ctx.v[0] = p.vsOut + (cpvk_u64)s0 * p.vsStride; ctx.v[1] = p.vsOut + (cpvk_u64)s1 * p.vsStride; ctx.v[2] = ctx.v[1]; ctx.vProv = ctx.v[0];
                        ctx.fragCoord[0] = cpvk_spec_u32(CPVK_SPEC_ORIGIN_UPPER) ? (float)x : p.vpWidth - (float)x - 1.0f;
                        ctx.fragCoord[1] = (float)y; ctx.fragCoord[2] = depth; ctx.fragCoord[3] = 1.0f;
                        ctx.dp = &p; ctx.unorm8 = sLut;
                        const float fragDepth = (p.vpMaxDepth - p.vpMinDepth) * depth + p.vpMinDepth;
                        if (!cpvk_fs_main(&ctx, &out)) wrote = rop(x - tileX0, y - tileY0, fragDepth, true, out);
                    }
                    nPass += __popc(__ballot_sync(0xFFFFFFFFu, wrote));
                }
            }
        }
    }
    #pragma unroll 1
    for (cpvk_u32 chunkBase = listBegin; triangles && chunkBase < listEnd; chunkBase += CPVK_CHUNK) {
        const int n = (int)min((cpvk_u32)CPVK_CHUNK, listEnd - chunkBase);
        if (!listsSorted && chunkBase == listBegin) {
            // Binning claims list slots with atomics, so a tile's ids arrive in arbitrary order. Lists of up to CPVK_ORDER_MAX ids
            // are ordered here, once, before the first chunk is staged (ids are unique; ascending id = API order); the host runs
            // k_bin_sort only for longer ones.
            // Usual case — the ids of a tile lie within CPVK_ORDER_BITS of each other (any draw of fewer primitives, any mesh
            // with some locality): every id sets its bit in a shared-memory bitmap, each thread counts the bits of its 128-bit
            // slice, a block-wide prefix sum gives the slice's first position, and the slices are written out in order.
            const int nAll = (int)(listEnd - listBegin);
            const uint4 mnA = reinterpret_cast<const uint4*>(sRange)[0], mnB = reinterpret_cast<const uint4*>(sRange)[1];
            const uint4 mxA = reinterpret_cast<const uint4*>(sRange)[2], mxB = reinterpret_cast<const uint4*>(sRange)[3];
            const cpvk_u32 lo = min(min(min(mnA.x, mnA.y), min(mnA.z, mnA.w)), min(min(mnB.x, mnB.y), min(mnB.z, mnB.w)));
            const cpvk_u32 hi = max(max(max(mxA.x, mxA.y), max(mxA.z, mxA.w)), max(max(mxB.x, mxB.y), max(mxB.z, mxB.w)));
            if (hi - lo < (cpvk_u32)CPVK_ORDER_BITS) { // block-uniform
                if (keyA != 0xFFFFFFFFu) atomicOr(sBits + ((keyA - lo) >> 5), 1u << ((keyA - lo) & 31u));
                if (keyB != 0xFFFFFFFFu) atomicOr(sBits + ((keyB - lo) >> 5), 1u << ((keyB - lo) & 31u));
                __syncthreads();
                const uint4 bits = reinterpret_cast<const uint4*>(sBits)[threadIdx.x];
                const cpvk_u32 cnt = (cpvk_u32)(__popc(bits.x) + __popc(bits.y)) + (cpvk_u32)(__popc(bits.z) + __popc(bits.w));
                cpvk_u32 incl = cnt;
                #pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const cpvk_u32 v = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += v; }
                if (lane == 31) sRange[16 + warp] = incl;
                __syncthreads();
                cpvk_u32 pos = incl - cnt;
                for (int q = 0; q < warp; q++) pos += sRange[16 + q];
                // the slices are written out by the whole warp, one non-empty 32-bit word at a time (its 32 ids = the 32 lanes):
                // a tile's ids come in a few dense runs, which a per-thread loop would leave to a few busy lanes
                const cpvk_u32 w4[4] = {bits.x, bits.y, bits.z, bits.w};
                #pragma unroll
                for (int q = 0; q < 4; q++) {
                    cpvk_u32 nonEmpty = __ballot_sync(0xFFFFFFFFu, w4[q] != 0u);
                    while (nonEmpty) { // warp-uniform
                        const int srcLane = __ffs((int)nonEmpty) - 1;
                        nonEmpty &= nonEmpty - 1u;
                        const cpvk_u32 word = __shfl_sync(0xFFFFFFFFu, w4[q], srcLane), first = __shfl_sync(0xFFFFFFFFu, pos, srcLane);
                        if ((word >> lane) & 1u)
                            sSorted[first + (cpvk_u32)__popc(word & ((1u << lane) - 1u))] = lo + ((cpvk_u32)(warp * 32 + srcLane) * 128u + (cpvk_u32)q * 32u) + (cpvk_u32)lane;
                    }
                    pos += (cpvk_u32)__popc(w4[q]);
                }
            } else {
                // ids too far apart for the bitmap: each id's rank (number of ids <= it, minus one) is its position. Broadcast
                // shared-memory reads, no barriers inside the loop.
                const cpvk_u32* sKeys = reinterpret_cast<const cpvk_u32*>(sBB);  // staged before the first barrier; reused before the bboxes are staged
                const uint4* k4 = reinterpret_cast<const uint4*>(sKeys);
                // rank += (v <= key): the carry of key - v (set when there is no borrow, i.e. v <= key) is added straight
                // into the rank, 1.5 instructions per pair. The key itself is counted once, hence the start value -1; the unused
                // slots hold 0xFFFFFFFF and count for no real key.
                #define CPVK_RANK_STEP(rk, v, key) asm("{ .reg .u32 t; sub.cc.u32 t, %2, %1; addc.u32 %0, %0, 0; }" : "+r"(rk) : "r"(v), "r"(key))
                cpvk_u32 rankA = 0xFFFFFFFFu, rankB = 0xFFFFFFFFu;
                for (int j = 0; j < (nAll + 3) / 4; j++) {
                    const uint4 v = k4[j];
                    CPVK_RANK_STEP(rankA, v.x, keyA); CPVK_RANK_STEP(rankA, v.y, keyA); CPVK_RANK_STEP(rankA, v.z, keyA); CPVK_RANK_STEP(rankA, v.w, keyA);
                    if (nAll > CPVK_CHUNK) { CPVK_RANK_STEP(rankB, v.x, keyB); CPVK_RANK_STEP(rankB, v.y, keyB); CPVK_RANK_STEP(rankB, v.z, keyB); CPVK_RANK_STEP(rankB, v.w, keyB); }
                }
                #undef CPVK_RANK_STEP
                if (keyA != 0xFFFFFFFFu) sSorted[rankA] = keyA;
                if (keyB != 0xFFFFFFFFu) sSorted[rankB] = keyB;
            }
            __syncthreads();
        }
        cpvk_u32 stagedPrim = 0;
        if ((int)threadIdx.x < n) stagedPrim = listsSorted ? __ldg(p.tileLists + chunkBase + threadIdx.x) : sSorted[(chunkBase - listBegin) + threadIdx.x];
        if ((int)threadIdx.x < n) { // CPVK_CHUNK == blockDim.x: one record per thread, 16-byte coalesced pieces
            const uint4* sp = reinterpret_cast<const uint4*>(p.setups + stagedPrim);
            #pragma unroll
            for (int j = 0; j < 6; j++) sQ[j * CPVK_CHUNK + threadIdx.x] = __ldg(sp + j);
            const uint2 b = __ldg(reinterpret_cast<const uint2*>(p.bboxes + stagedPrim));
            sBB[threadIdx.x] = b;
            // which of the eight 16x8 warp regions the bbox meets (bit = warp index), worked out once per triangle here
            // instead of once per triangle per warp in the compaction scan below
            const int bx0 = (short)(b.x & 0xFFFFu), by0 = (short)(b.x >> 16), bx1 = (short)(b.y & 0xFFFFu), by1 = (short)(b.y >> 16);
            cpvk_u32 xb = 0, m = 0;
            #pragma unroll
            for (int wx = 0; wx < 2; wx++) { const int a = tileX0 + wx * CPVK_REGION_W; if (bx0 < min(a + CPVK_REGION_W, x1) && bx1 > a) xb |= 1u << wx; }
            #pragma unroll
            for (int wy = 0; wy < 4; wy++) { const int a = tileY0 + wy * CPVK_REGION_H; if (by0 < min(a + CPVK_REGION_H, y1) && by1 > a) m |= xb << (2 * wy); }
            sMask[threadIdx.x] = (cpvk_u8)m;
        }
        __syncthreads();
        if (regionLive) {
            // ---- step 1: compact the triangles whose bbox meets this warp's region (order kept) ----
            int nHits = 0;
            #pragma unroll 1
            for (int base = 0; base < n; base += 32) {
                const int li = base + lane;
                const bool hit = li < n && ((sMask[li] >> warp) & 1u) != 0u;
                const cpvk_u32 m = __ballot_sync(0xFFFFFFFFu, hit);
                if (hit) sHit[nHits + __popc(m & ((1u << lane) - 1u))] = (cpvk_u8)li;
                nHits += __popc(m);
            }
            __syncwarp();
            // ---- step 2: 32 hit triangles at a time, one per lane ----
            // Small triangles (<= 32 candidate pixels in this region) get their coverage mask computed by their own lane;
            // their covered pixels are then packed densely, 32 fragments per shading batch, in triangle order. A large
            // triangle is rasterised by the whole warp, one pixel per lane, and splits the packing at its position so
            // that fragments still reach the ROP in API order.
            #pragma unroll 1
            for (int hb = 0; hb < nHits; hb += 32) {
                const bool valid = hb + lane < nHits;
                const cpvk_u32 kt = valid ? sHit[hb + lane] : 0u;
                int cx0 = 0, cy0 = 0, cw = 0, ch = 0;
                if (valid) {
                    const uint2 b = sBB[kt];
                    const int bx0 = (short)(b.x & 0xFFFFu), by0 = (short)(b.x >> 16), bx1 = (short)(b.y & 0xFFFFu), by1 = (short)(b.y >> 16);
                    cx0 = max(bx0, rx0); cy0 = max(by0, ry0);
                    cw = min(bx1, rx1) - cx0; ch = min(by1, ry1) - cy0;
                }
                const int cand = cw * ch; // 1 .. 128
                const bool small = valid && cand <= 32;
                cpvk_u32 cov = 0; // bit c = candidate c (row-major inside the candidate rectangle) is covered
                {
                    float e0ax = 0, e0ay = 0, e0dy = 0, e0dx = 0, e1ax = 0, e1ay = 0, e1dy = 0, e1dx = 0, e2ax = 0, e2ay = 0, e2dy = 0, e2dx = 0;
                    if (small) {
                        const uint4 q0 = sQ[kt], q1 = sQ[CPVK_CHUNK + kt], q2 = sQ[2 * CPVK_CHUNK + kt];
                        e0ax = __uint_as_float(q0.x); e0ay = __uint_as_float(q0.y); e0dy = __uint_as_float(q0.z); e0dx = __uint_as_float(q0.w);
                        e1ax = __uint_as_float(q1.x); e1ay = __uint_as_float(q1.y); e1dy = __uint_as_float(q1.z); e1dx = __uint_as_float(q1.w);
                        e2ax = __uint_as_float(q2.x); e2ay = __uint_as_float(q2.y); e2dy = __uint_as_float(q2.z); e2dx = __uint_as_float(q2.w);
                    }
                    // EdgeFunction at the pixel centre (Draw.cpp:415-418); inside = none of the three is < 0 (no fill rule,
                    // NaN accepted). w = A - B with A = (x - ax) * dy, B = (y - ay) * dx, and fl(A - B) < 0 exactly when
                    // A < B, so the coverage test compares the two products and skips the subtraction.
                    const int maxCand = __reduce_max_sync(0xFFFFFFFFu, small ? cand : 0);
                    const int maxW = __reduce_max_sync(0xFFFFFFFFu, small ? cw : 0), maxH = __reduce_max_sync(0xFFFFFFFFu, small ? ch : 0);
                    const int bx = small ? cx0 - tileX0 : 0, by = small ? cy0 - tileY0 : 0;
                    if (maxW * maxH * 5 <= maxCand * 8) {
                        // similar rectangles across the lanes: all lanes walk one maxW x maxH window in lockstep, which
                        // hoists the row terms out of the pixel loop. A row's results are collected at the warp-uniform bit
                        // xx, trimmed to the lane's own width and appended at yy * cw, so candidate numbering stays
                        // row-major per lane.
                        const cpvk_u32 rowMask = small ? (cw >= 32 ? 0xFFFFFFFFu : ((1u << cw) - 1u)) : 0u; // cw <= CPVK_REGION_W <= 32
                        const int rows = small ? ch : 0;
                        int shift = 0;
                        // CPVK_COVER_ROWS rows per pass: A_k(x) = (xf - ax_k) * dy_k does not depend on the row and B_k(y) = (yf - ay_k) *
                        // dx_k not on the column, so a column's three A terms are compared against both rows' B terms
                        // (same operations on the same operands as the per-pixel expression: identical bits).
                        const float* yp = sYf + by; // by + yy + CPVK_COVER_ROWS - 1 < 48: stays inside the xf/yf/lut block, masked by `rows`
                        #pragma unroll 1
                        for (int yy = 0; yy < maxH; yy += CPVK_COVER_ROWS) {
                            float b0[CPVK_COVER_ROWS], b1[CPVK_COVER_ROWS], b2[CPVK_COVER_ROWS];
                            cpvk_u32 r[CPVK_COVER_ROWS];
                            #pragma unroll
                            for (int k = 0; k < CPVK_COVER_ROWS; k++) {
                                const float yf = yp[yy + k];
                                b0[k] = (yf - e0ay) * e0dx; b1[k] = (yf - e1ay) * e1dx; b2[k] = (yf - e2ay) * e2dx; r[k] = 0;
                            }
                            cpvk_u32 bit = 1u;
                            const float* xp = sXf + bx; // bx + xx <= 46: stays inside the xf/yf/lut block
                            const float* const xe = xp + maxW;
                            constexpr int kCoverUnroll = CPVK_COVER_UNROLL;
                            #pragma unroll kCoverUnroll
                            for (; xp != xe; xp++, bit <<= 1) {
                                const float xf = *xp;
                                const float a0 = (xf - e0ax) * e0dy, a1 = (xf - e1ax) * e1dy, a2 = (xf - e2ax) * e2dy;
                                // inside = !(a0 < b0 || a1 < b1 || a2 < b2) = (a0 >=u b0) && (a1 >=u b1) && (a2 >=u b2), "u" = or unordered:
                                // three chained predicate compares and one predicated OR per pixel
                                #pragma unroll
                                for (int k = 0; k < CPVK_COVER_ROWS; k++)
                                    asm("{ .reg .pred q; setp.geu.f32 q, %1, %2; setp.geu.and.f32 q, %3, %4, q; setp.geu.and.f32 q, %5, %6, q; @q or.b32 %0, %0, %7; }"
                                        : "+r"(r[k]) : "f"(a0), "f"(b0[k]), "f"(a1), "f"(b1[k]), "f"(a2), "f"(b2[k]), "r"(bit));
                            }
                            #pragma unroll
                            for (int k = 0; k < CPVK_COVER_ROWS; k++)
                                if (yy + k < rows) { cov |= (r[k] & rowMask) << shift; shift += cw; }
                        }
                    } else {
                        int xx = 0, yy = 0;
                        #pragma unroll 1
                        for (int c = 0; c < maxCand; c++) {
                            if (small && c < cand) {
                                const float xf = sXf[bx + xx], yf = sYf[by + yy];
                                const bool out = (xf - e0ax) * e0dy < (yf - e0ay) * e0dx || (xf - e1ax) * e1dy < (yf - e1ay) * e1dx ||
                                                 (xf - e2ax) * e2dy < (yf - e2ay) * e2dx;
                                if (!out) cov |= 1u << c;
                                if (++xx == cw) { xx = 0; yy++; }
                            }
                        }
                    }
                }
                // A large triangle whose three edge functions cannot all be >= 0 anywhere in its candidate rectangle is dropped
                // before the warp walks it (a full-screen quad's second triangle misses half of the regions its bbox meets).
                // Exact, not conservative-by-epsilon: w = fl(A(x) - B(y)) with A(x) = fl(fl(xf - ax) * dy), B(y) = fl(fl(yf - ay)
                // * dx); xf, yf are monotone in x, y and every rounded operation is monotone, so over the rectangle w is at
                // most fl(max(A(x_first), A(x_last)) - min(B(y_first), B(y_last))) — if that is < 0, every pixel fails this
                // edge's `w < 0` test. Only taken when the edge constants are finite and small enough that no product can
                // overflow (otherwise an inf * 0 = NaN in the interior, which the reference accepts, could hide from the corners).
                bool large = valid && !small;
                if (__any_sync(0xFFFFFFFFu, large)) {
                    if (large) {
                        const float xa = sXf[cx0 - tileX0], xb = sXf[cx0 + cw - 1 - tileX0], ya = sYf[cy0 - tileY0], yb = sYf[cy0 + ch - 1 - tileY0];
                        bool reject = false;
                        #pragma unroll
                        for (int e = 0; e < 3; e++) {
                            const uint4 q = sQ[e * CPVK_CHUNK + kt];
                            const float ax = __uint_as_float(q.x), ay = __uint_as_float(q.y), dy = __uint_as_float(q.z), dx = __uint_as_float(q.w);
                            const bool bounded = fabsf(ax) <= 1e15f && fabsf(ay) <= 1e15f && fabsf(dy) <= 1e15f && fabsf(dx) <= 1e15f;
                            const float aMax = fmaxf((xa - ax) * dy, (xb - ax) * dy), bMin = fminf((ya - ay) * dx, (yb - ay) * dx);
                            reject = reject || (bounded && aMax < bMin);
                        }
                        large = !reject;
                    }
                }
                const cpvk_u32 largeMask = __ballot_sync(0xFFFFFFFFu, large);
                cpvk_u32 todo = __ballot_sync(0xFFFFFFFFu, small || large);
                // Covered candidates of the small triangles are written, triangle after triangle (= API order), to this
                // warp's fragment list in shared memory: 16 bits each = chunk-local triangle | x, y inside the warp region.
                // The list is then shaded 32 fragments at a time. A large triangle ends the segment and is rasterised by
                // the whole warp right after it, so fragments still reach the ROP in API order.
                const cpvk_u32 divMagic = 1024u / (cpvk_u32)max(cw, 1) + 1u; // (c * divMagic) >> 10 == c / cw for c < 32, cw <= 32
                const cpvk_u32 fragBase = kt | ((cpvk_u32)(cx0 - rx0) << 8) | ((cpvk_u32)(cy0 - ry0) << 13); // triangle (8 bits) | x in the region (5) | y in the region (3)
                #pragma unroll 1
                while (todo) {
                    const cpvk_u32 lt = largeMask & todo;
                    int firstLarge = lt ? __ffs(lt) - 1 : 32;
                    cpvk_u32 seg = firstLarge == 32 ? todo : (todo & ((1u << firstLarge) - 1u));
                    int cnt = ((seg >> lane) & 1u) ? __popc(cov) : 0;
                    int incl = cnt;
                    #pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += v; }
                    if (__shfl_sync(0xFFFFFFFFu, incl, 31) > CPVK_FRAG_CAP) {
                        // more fragments than the list holds: keep the longest prefix that fits, the rest (and the large
                        // triangle, which must come after them) waits for the next round
                        seg &= __ballot_sync(0xFFFFFFFFu, incl <= CPVK_FRAG_CAP);
                        if (!((seg >> lane) & 1u)) cnt = 0;
                        firstLarge = 32;
                    }
                    const int total = (int)__reduce_max_sync(0xFFFFFFFFu, (unsigned)(cnt ? incl : 0));
                    {
                        int pos = incl - cnt;
                        cpvk_u32 m = cnt ? cov : 0u;
                        while (m) {
                            const cpvk_u32 c = (cpvk_u32)__ffs((int)m) - 1u;
                            m &= m - 1u;
                            const cpvk_u32 row = (c * divMagic) >> 10;
                            sFrag[pos++] = (unsigned short)(fragBase + ((c - row * (cpvk_u32)cw) << 8) + (row << 13));
                        }
                    }
                    __syncwarp();
                    // large triangle (warp-uniform)
                    int lcx0 = 0, lcy0 = 0, lcx1 = 0, lcy1 = 0, lg = 4; cpvk_u32 lkt = 0;
                    if (firstLarge < 32) {
                        lkt = __shfl_sync(0xFFFFFFFFu, kt, firstLarge);
                        lcx0 = __shfl_sync(0xFFFFFFFFu, cx0, firstLarge); lcy0 = __shfl_sync(0xFFFFFFFFu, cy0, firstLarge);
                        lcx1 = lcx0 + __shfl_sync(0xFFFFFFFFu, cw, firstLarge); lcy1 = lcy0 + __shfl_sync(0xFFFFFFFFu, ch, firstLarge);
                        const int lw = lcx1 - lcx0;
                        lg = lw <= 4 ? 2 : (lw <= 8 ? 3 : (lw <= 16 ? 4 : 5)); // lanes form a (1<<lg) x (32>>lg) block of candidates
                    }
                    int o = 0, row0 = lcy0;
                    #pragma unroll 1
                    for (;;) {
                        bool active = false, distinct = false; cpvk_u32 bkt = 0; int px = 0, py = 0;
                        if (o < total) {
                            active = o + lane < total;
                            if (active) {
                                const cpvk_u32 rec = sFrag[o + lane];
                                bkt = rec & 255u; px = rx0 - tileX0 + (int)((rec >> 8) & 31u); py = ry0 - tileY0 + (int)(rec >> 13);
                            }
                            o += 32;
                        } else if (firstLarge < 32 && row0 < lcy1) {
                            const int x = lcx0 + (lane & ((1 << lg) - 1)), y = row0 + (lane >> lg);
                            row0 += 32 >> lg;
                            active = x < lcx1 && y < lcy1; bkt = lkt; px = x - tileX0; py = y - tileY0; distinct = true; // coverage is tested by shadeBatch
                        } else break;
                        shadeBatch(active, bkt, px, py, distinct); // the only call site: one copy of the fragment shader per kernel
                    }
                    __syncwarp();
                    todo &= ~seg;
                    if (firstLarge < 32) todo &= ~(1u << firstLarge);
                }
            }
        }
        if (chunkBase + CPVK_CHUNK < listEnd) __syncthreads(); // the staged chunk is free for the next one (after the last chunk the barrier below does it)
    }
#ifndef CPVK_WARP_WRITEBACK
#define CPVK_WARP_WRITEBACK 0 /* 1: every warp stores its own region without the tile-wide barrier — measured 1 % slower at C3/M1 (64-byte row segments instead of 128) */
#endif
#if CPVK_WARP_WRITEBACK
    // ---- write the tile back: shared -> HBM. A warp's region was written by that warp alone, so each warp could store its own region as
    // soon as it is done with the tile's last chunk — no barrier, nobody waits for the slowest warp of the tile (tuning variant). ----
    __syncwarp();
    {
        const int wy = max(ry0, wy0); // the band's first row may cut the region
        if (rx0 < rx1 && wy < ry1) {
            const cpvk_u32 ox = (cpvk_u32)(rx0 - tileX0), oy = (cpvk_u32)(wy - tileY0), rw = (cpvk_u32)(rx1 - rx0), rh = (cpvk_u32)(ry1 - wy);
            if (dsUsed && ((depthTest && depthWrite) || stencilOn || (lazyMask & 0x100u)))
                cpvk_region_copy(reinterpret_cast<cpvk_u8*>(p.ds.address) + (cpvk_u64)wy * p.ds.rowPitch + (cpvk_u64)rx0 * dsTexel, p.ds.rowPitch,
                                 sDepth + oy * dsPitch + ox * dsTexel, dsPitch, rw * dsTexel, rh);
            #pragma unroll
            for (int a = 0; a < CPVK_MAX_COLOR; a++)
                if (sColor[a])
                    cpvk_region_copy(reinterpret_cast<cpvk_u8*>(p.color[a].address) + (cpvk_u64)wy * p.color[a].rowPitch + (cpvk_u64)rx0 * cTexel[a], p.color[a].rowPitch,
                                     sColor[a] + (oy * CPVK_TILE_W + ox) * cTexel[a], cTexel[a] * CPVK_TILE_W, rw * cTexel[a], rh);
            // ---- the fused gather: the band's rows of this region go to every peer's copy of colour attachment 0 ----
            if (p.mirrorCount && sColor[0])
                for (cpvk_u32 m = 0; m < p.mirrorCount; m++)
                    cpvk_region_copy(reinterpret_cast<cpvk_u8*>(p.mirror[m]) + (cpvk_u64)wy * p.color[0].rowPitch + (cpvk_u64)rx0 * cTexel[0], p.color[0].rowPitch,
                                     sColor[0] + (oy * CPVK_TILE_W + ox) * cTexel[0], cTexel[0] * CPVK_TILE_W, rw * cTexel[0], rh);
        }
    }
#else
    __syncthreads();
    // ---- write the tile back: shared -> HBM, row segments are contiguous in the linear image ----
    if (dsUsed && ((depthTest && depthWrite) || stencilOn || (lazyMask & 0x100u)))
        cpvk_tile_copy(reinterpret_cast<cpvk_u8*>(p.ds.address) + (cpvk_u64)wy0 * p.ds.rowPitch + (cpvk_u64)tileX0 * dsTexel, p.ds.rowPitch,
                       sDepth + skipRows * dsPitch, dsPitch, (cpvk_u32)tw * dsTexel, (cpvk_u32)th, dsPitch);
    #pragma unroll
    for (int a = 0; a < CPVK_MAX_COLOR; a++)
        if (sColor[a])
            cpvk_tile_copy(reinterpret_cast<cpvk_u8*>(p.color[a].address) + (cpvk_u64)wy0 * p.color[a].rowPitch + (cpvk_u64)tileX0 * cTexel[a], p.color[a].rowPitch,
                           sColor[a] + skipRows * cTexel[a] * CPVK_TILE_W, cTexel[a] * CPVK_TILE_W, (cpvk_u32)tw * cTexel[a], (cpvk_u32)th, cTexel[a] * CPVK_TILE_W);
    // ---- the fused gather: the band's rows of this tile go to every peer's copy of colour attachment 0 ----
    if (p.mirrorCount && sColor[0])
        for (cpvk_u32 m = 0; m < p.mirrorCount; m++)
            cpvk_tile_copy(reinterpret_cast<cpvk_u8*>(p.mirror[m]) + (cpvk_u64)wy0 * p.color[0].rowPitch + (cpvk_u64)tileX0 * cTexel[0], p.color[0].rowPitch,
                           sColor[0] + skipRows * cTexel[0] * CPVK_TILE_W, cTexel[0] * CPVK_TILE_W, (cpvk_u32)tw * cTexel[0], (cpvk_u32)th, cTexel[0] * CPVK_TILE_W);
#endif
    if (p.stats && triangles && threadIdx.x == 0 && listEnd != listBegin) atomicAdd(p.stats + 2, (cpvk_u64)(listEnd - listBegin));
    if (p.stats && lane == 0 && (nCov | nPass)) {
        atomicAdd(p.stats + 0, (cpvk_u64)nCov);
        atomicAdd(p.stats + 1, (cpvk_u64)nPass);
    }
}
