// fixed_kernels.cu — sm_100a kernels that do not depend on application shaders: triangle setup, API-order
// preserving tile binning, clear, row copy and blit. Compiled straight to SASS into libcpvk_cuda.so.
//
//   k_setup      a2 + a5 (per-primitive part) + a6 (area, facing, cull, bbox) of SURVEY §8(a)
//   k_bin_*      new: screen-tile lists that keep API primitive order (count -> scan -> fill -> per-tile sort)
//   k_clear      a15 / f1: ClearImage (Draw.cpp:117-149) as vector stores of one packed texel
//   k_copy_rows  f2: vkCmdCopyImage row memcpy (CommandBuffer.Copy.cpp:77-200)
//   k_blit       f2: vkCmdBlitImage (CommandBuffer.cpp:57-232)
#include <cstdint>

#include "cpvk_device.cuh"
#include "fixed_kernels.h"

// x86 cvttss2si semantics for static_cast<int32_t>(float) as the reference binary executes it
// (Draw.cpp:1548-1564): NaN and out-of-range inputs give INT_MIN ("integer indefinite").
__device__ __forceinline__ int cpvk_cvtt(float v) {
    return (v >= 2147483648.0f || v < -2147483648.0f || v != v) ? (int)0x80000000 : (int)v;
}

// Warp-aggregated walk over the (at most CPVK_BIN_SMALL) tiles of each lane's bbox: lanes that name the same tile in
// the same step are served by one atomic. Consecutive primitives mostly land in the same few tiles, so this cuts
// the same-address atomics by an order of magnitude. fn(tile, leader, lanesOnThisTile, rankAmongThem).
template <typename F> __device__ __forceinline__ void cpvk_for_each_tile_aggregated(bool has, int tx0, int ty0, int tw, int n, cpvk_u32 tilesX, F fn) {
    const int lane = threadIdx.x & 31;
    const int maxN = (int)__reduce_max_sync(0xFFFFFFFFu, (unsigned)(has ? n : 0));
    int cx = 0, cy = 0;
    for (int j = 0; j < maxN; j++) {
        const bool live = has && j < n;
        const cpvk_u32 t = live ? (cpvk_u32)(ty0 + cy) * tilesX + (cpvk_u32)(tx0 + cx) : 0xFFFFFFFFu;
        const cpvk_u32 m = __match_any_sync(0xFFFFFFFFu, t);
        fn(t, live, m, lane);
        if (++cx == tw) { cx = 0; cy++; }
    }
}

// The last CTA of a grid to get here stores the verdict of single-pass binning where the host reads it (mapped memory):
// [2] deferred primitives, [3] != 0 = replay with exact sizes (a tile list overflowed, or deferred primitives turned up in
// a draw planned without a k_bin_large pass). Totals are not known in this mode: [0] = [1] = 0.
// Ordering: meta[2] only changes through atomics whose value the thread uses (performed before it goes on); a thread that
// raises meta[3] fences right there (cpvk_raise_mismatch, rare), before the barrier below — so the ticket needs no fence of
// its own, which measured 7 us of k_setup on the 1M-triangle draw.
__device__ __forceinline__ void cpvk_raise_mismatch(cpvk_u32* meta) { *(volatile cpvk_u32*)(meta + 3) = 1u; __threadfence(); }
__device__ __forceinline__ void cpvk_publish_when_last(cpvk_u32* ticket, cpvk_u32* meta, cpvk_u32* metaHost, bool largeDone) {
    __syncthreads();
    if (threadIdx.x != 0) return;
    if (atomicAdd(ticket, 1u) != gridDim.x - 1) return;
    const cpvk_u32 nLarge = *(volatile cpvk_u32*)(meta + 2);
    cpvk_u32 mismatch = *(volatile cpvk_u32*)(meta + 3);
    if (nLarge != 0 && !largeDone) { mismatch = 1u; *(volatile cpvk_u32*)(meta + 3) = 1u; }
    metaHost[0] = 0u; metaHost[1] = 0u; metaHost[2] = nLarge; metaHost[3] = mismatch;
    __threadfence_system();
}

__global__ void __launch_bounds__(256) k_setup(CpvkSetupArgs a) {
    __shared__ uint4 sRec[256 / 32][32 * 6]; // one warp's 32 records, staged so that the HBM stores are contiguous 512-byte rows
    const cpvk_u32 pRaw = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = pRaw < a.primCount;
    const cpvk_u32 p = valid ? pRaw : a.primCount - 1; // idle lanes of the last warp redo the last primitive, results dropped
    // CalculatePrimitives (Draw.cpp:614-661)
    cpvk_u32 i0, i1, i2, prov;
    if (a.topology == 3) { prov = p * 3; i0 = p * 3; i1 = p * 3 + 1; i2 = p * 3 + 2; }
    else if (a.topology == 4) { prov = p; i0 = p; i1 = p + 1; i2 = p + 2; }
    else { prov = p + 1; i0 = 0; i1 = p + 1; i2 = p + 2; }
    if (a.frontFace == 1) { const cpvk_u32 t = i0; i0 = i2; i2 = t; } // Draw.cpp:1532-1535
    if (cpvk_vcache_on(a.vcache, a.nVerts)) { // stream position -> shaded vertex
        const cpvk_u32 lo = cpvk_vcache_lowest(a.vcache);
        i0 = cpvk_fetch_index(a.indexBuffer, a.indexStride, (cpvk_u64)a.first + i0) - lo;
        i1 = cpvk_fetch_index(a.indexBuffer, a.indexStride, (cpvk_u64)a.first + i1) - lo;
        i2 = cpvk_fetch_index(a.indexBuffer, a.indexStride, (cpvk_u64)a.first + i2) - lo;
        prov = cpvk_fetch_index(a.indexBuffer, a.indexStride, (cpvk_u64)a.first + prov) - lo;
    }
    const cpvk_u32 idx[3] = {i0, i1, i2};
    float P[3][4];
    #pragma unroll
    for (int k = 0; k < 3; k++) {
        const uint4 pv = __ldg(a.vsPos + idx[k]); // already position / position.w with .w = position.w (cpvk_store_position; Draw.cpp:1541-1546)
        P[k][0] = __uint_as_float(pv.x); P[k][1] = __uint_as_float(pv.y); P[k][2] = __uint_as_float(pv.z); P[k][3] = __uint_as_float(pv.w);
    }
    const float W = a.vpWidth, H = a.vpHeight;
    int sx[3], sy[3];
    #pragma unroll
    for (int k = 0; k < 3; k++) {
        sx[k] = cpvk_cvtt((P[k][0] + 1.0f) * 0.5f * W);
        sy[k] = cpvk_cvtt((P[k][1] + 1.0f) * 0.5f * H);
    }
    int startX = max(0, min(sx[0], min(sx[1], sx[2])));
    int startY = max(0, min(sy[0], min(sy[1], sy[2])));
    int endX = min(cpvk_cvtt(W), max(sx[0], max(sx[1], sx[2])) + 1);
    int endY = min(cpvk_cvtt(H), max(sy[0], max(sy[1], sy[2])) + 1);
    // render area: attachments and this GPU's band (the reference has neither clamp: writes past the attachment
    // are undefined behaviour there)
    startX = max(startX, a.clipX0); startY = max(startY, a.clipY0);
    endX = min(endX, a.clipX1); endY = min(endY, a.clipY1);

    // GetFragmentInput (Draw.cpp:879-903): area, facing, edge orientation, cull
    float area = (P[2][0] - P[0][0]) * (P[1][1] - P[0][1]) - (P[2][1] - P[0][1]) * (P[1][0] - P[0][0]);
    bool front;
    int ea[3], eb[3];
    if (area < 0.0f) { area = -area; front = false; ea[0] = 2; eb[0] = 1; ea[1] = 0; eb[1] = 2; ea[2] = 1; eb[2] = 0; }
    else { front = true; ea[0] = 1; eb[0] = 2; ea[1] = 2; eb[1] = 0; ea[2] = 0; eb[2] = 1; }
    const bool culled = ((a.cullMode & 2u) && !front) || ((a.cullMode & 1u) && front);

    CpvkTriSetup s;
    #pragma unroll
    for (int k = 0; k < 3; k++) {
        float ax = 0, ay = 0, bx = 0, by = 0;
        #pragma unroll
        for (int j = 0; j < 3; j++) { if (ea[k] == j) { ax = P[j][0]; ay = P[j][1]; } if (eb[k] == j) { bx = P[j][0]; by = P[j][1]; } }
        s.e[k][0] = ax; s.e[k][1] = ay; s.e[k][2] = by - ay; s.e[k][3] = bx - ax;
        s.z[k] = P[k][2]; s.pw[k] = P[k][3]; s.idx[k] = idx[k];
    }
    s.area = area; s.flags = front ? 1u : 0u; s.provoking = prov;
    // A primitive that is culled or outside the render area (this GPU's band) is never listed by a tile, so its record
    // is never read: a warp whose 32 primitives are all like that skips the 3 KB store. With N bands that is most warps.
    const bool listed = valid && !culled && endX > startX && endY > startY;
    if (__any_sync(0xFFFFFFFFu, listed)) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const uint4* sv = reinterpret_cast<const uint4*>(&s);
        #pragma unroll
        for (int j = 0; j < 6; j++) sRec[warp][lane * 6 + j] = sv[j];
        __syncwarp();
        const cpvk_u32 warpFirst = pRaw - (cpvk_u32)lane;
        const cpvk_u32 nRec = warpFirst < a.primCount ? min(32u, a.primCount - warpFirst) : 0u;
        uint4* dst = reinterpret_cast<uint4*>(a.setups + warpFirst);
        #pragma unroll
        for (int j = 0; j < 6; j++) if ((cpvk_u32)(j * 32 + lane) < nRec * 6u) dst[j * 32 + lane] = sRec[warp][j * 32 + lane];
    }
    CpvkBBox bb;
    if (!valid || culled || endX <= startX || endY <= startY) { bb.x0 = bb.y0 = bb.x1 = bb.y1 = 0; }
    else { bb.x0 = (short)startX; bb.y0 = (short)startY; bb.x1 = (short)endX; bb.y1 = (short)endY; }
    if (valid) a.bboxes[p] = bb;
    // binning pass 0, fused: count the (primitive, tile) pairs while the bbox is in registers
    bool small = false; int tx0 = 0, ty0 = 0, tw = 1, n = 0;
    if (bb.x1 > bb.x0) {
        tx0 = bb.x0 / CPVK_TILE_W; ty0 = bb.y0 / CPVK_TILE_H - (int)a.tileRow0;
        const int tx1 = (bb.x1 - 1) / CPVK_TILE_W, ty1 = (bb.y1 - 1) / CPVK_TILE_H - (int)a.tileRow0;
        tw = tx1 - tx0 + 1; n = tw * (ty1 - ty0 + 1);
        if (n > CPVK_BIN_SMALL) a.largeList[atomicAdd(a.meta + 2, 1u)] = p;
        else small = true;
    }
    if (!a.directLists) {
        cpvk_for_each_tile_aggregated(small, tx0, ty0, tw, n, a.tilesX, [&](cpvk_u32 t, bool live, cpvk_u32 m, int lane) {
            if (live && (int)__ffs((int)m) - 1 == lane) atomicAdd(a.counts + t, (cpvk_u32)__popc(m));
        });
        return;
    }
    // single-pass binning: the counting atomic also claims the slots, ids ascending inside one claim; k_raster orders
    // each list (at most directCap <= CPVK_ORDER_MAX ids), which restores API order exactly as after count -> scan -> fill
    bool overflow = false;
    {
        // four steps of the tile walk at a time: their claims are in flight together (an atomic that returns a value costs a
        // round trip to L2; one per step in a dependent chain was the whole cost of this mode), then the ids are stored
        const int lane = threadIdx.x & 31;
        const int maxN = (int)__reduce_max_sync(0xFFFFFFFFu, (unsigned)(small ? n : 0));
        int cx = 0, cy = 0;
        for (int j0 = 0; j0 < maxN; j0 += 4) {
            cpvk_u32 t[4], m[4], base[4];
            #pragma unroll
            for (int u = 0; u < 4; u++) {
                t[u] = 0xFFFFFFFFu; m[u] = 0u; base[u] = 0u;
                if (j0 + u < maxN) { // uniform
                    if (small && j0 + u < n) {
                        t[u] = (cpvk_u32)(ty0 + cy) * a.tilesX + (cpvk_u32)(tx0 + cx);
                        if (++cx == tw) { cx = 0; cy++; }
                    }
                    m[u] = __match_any_sync(0xFFFFFFFFu, t[u]);
                    if (t[u] != 0xFFFFFFFFu && (int)__ffs((int)m[u]) - 1 == lane) base[u] = atomicAdd(a.counts + t[u], (cpvk_u32)__popc(m[u]));
                }
            }
            #pragma unroll
            for (int u = 0; u < 4; u++) {
                if (j0 + u < maxN) {
                    const cpvk_u32 b = __shfl_sync(0xFFFFFFFFu, base[u], (int)__ffs((int)m[u]) - 1);
                    if (t[u] != 0xFFFFFFFFu) {
                        const cpvk_u32 slot = b + (cpvk_u32)__popc(m[u] & ((1u << lane) - 1u));
                        if (slot < a.directCap) a.directLists[(size_t)t[u] * a.directCap + slot] = p; else overflow = true;
                    }
                }
            }
        }
    }
    if (overflow) cpvk_raise_mismatch(a.meta);
    if (a.publish) cpvk_publish_when_last(a.ticket, a.meta, a.metaHost, false);
}

// Lowest and highest index of an indexed draw (vertex reuse, CpvkDrawParams::vcache).
__global__ void __launch_bounds__(256) k_index_range(cpvk_u64 indexBuffer, cpvk_u32 indexStride, cpvk_u32 first, cpvk_u32 count, cpvk_u32* range) {
    cpvk_u32 lo = 0xFFFFFFFFu, hi = 0u;
    for (cpvk_u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const cpvk_u32 v = cpvk_fetch_index(indexBuffer, indexStride, (cpvk_u64)first + i);
        lo = min(lo, v); hi = max(hi, v);
    }
    lo = __reduce_min_sync(0xFFFFFFFFu, lo); hi = __reduce_max_sync(0xFFFFFFFFu, hi);
    if ((threadIdx.x & 31) == 0) { atomicMax(range, ~lo); atomicMax(range + 1, hi); } // the lowest index is kept complemented so that a zero fill initialises both
}

// ---- binning ----
// Pass 0 counts (primitive, tile) pairs, pass 1 writes primitive ids at atomically claimed positions inside each
// tile's segment; k_bin_sort then sorts every segment ascending, which restores API order exactly (ids are unique).
// Primitives touching more than CPVK_BIN_SMALL tiles are deferred to k_bin_large, one CTA per primitive.
__global__ void __launch_bounds__(256) k_bin(CpvkBinArgs a, int pass) {
    if (pass != 0 && a.meta[3] != 0) return; // plan mismatch: the host replays the fill
    const cpvk_u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    bool small = false; int tx0 = 0, ty0 = 0, tw = 1, n = 0;
    if (p < a.primCount) {
        const CpvkBBox bb = a.bboxes[p];
        if (bb.x1 > bb.x0) {
            tx0 = bb.x0 / CPVK_TILE_W; ty0 = bb.y0 / CPVK_TILE_H - (int)a.tileRow0;
            const int tx1 = (bb.x1 - 1) / CPVK_TILE_W, ty1 = (bb.y1 - 1) / CPVK_TILE_H - (int)a.tileRow0;
            tw = tx1 - tx0 + 1; n = tw * (ty1 - ty0 + 1);
            small = n <= CPVK_BIN_SMALL; // larger ones are deferred to k_bin_large (listed by k_setup)
        }
    }
    cpvk_for_each_tile_aggregated(small, tx0, ty0, tw, n, a.tilesX, [&](cpvk_u32 t, bool live, cpvk_u32 m, int lane) {
        const int leader = (int)__ffs((int)m) - 1;
        if (pass == 0) { if (live && leader == lane) atomicAdd(a.counts + t, (cpvk_u32)__popc(m)); return; }
        cpvk_u32 base = 0;
        if (live && leader == lane) base = atomicAdd(a.cursors + t, (cpvk_u32)__popc(m));
        base = __shfl_sync(0xFFFFFFFFu, base, leader);
        if (live) a.lists[base + (cpvk_u32)__popc(m & ((1u << lane) - 1u))] = p; // ascending primitive id inside the claim
    });
}
__global__ void __launch_bounds__(256) k_bin_large(CpvkBinArgs a, int pass) {
    if (pass == 1 && a.meta[3] != 0) return;
    const cpvk_u32 nLarge = a.meta[2];
    for (cpvk_u32 li = blockIdx.x; li < nLarge; li += gridDim.x) {
        const cpvk_u32 p = a.largeList[li];
        const CpvkBBox bb = a.bboxes[p];
        const int tx0 = bb.x0 / CPVK_TILE_W, tx1 = (bb.x1 - 1) / CPVK_TILE_W, ty0 = bb.y0 / CPVK_TILE_H - (int)a.tileRow0, ty1 = (bb.y1 - 1) / CPVK_TILE_H - (int)a.tileRow0;
        const int tw = tx1 - tx0 + 1, n = tw * (ty1 - ty0 + 1);
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            const int ty = ty0 + k / tw, tx = tx0 + k % tw;
            const cpvk_u32 t = (cpvk_u32)ty * a.tilesX + (cpvk_u32)tx;
            if (pass == 0) atomicAdd(a.counts + t, 1u);
            else if (pass == 1) a.lists[atomicAdd(a.cursors + t, 1u)] = p;
            else {
                const cpvk_u32 slot = atomicAdd(a.counts + t, 1u);
                if (slot < a.directCap) a.directLists[(size_t)t * a.directCap + slot] = p; else cpvk_raise_mismatch(a.meta);
            }
        }
    }
    if (pass == 2) cpvk_publish_when_last(a.ticket, a.meta, a.metaHost, true);
}
// Single-CTA exclusive scan of the per-tile counts (<= a few 10^4 tiles). offsets[tiles] = total.
// meta[0] = total entries, meta[1] = longest list.
__global__ void __launch_bounds__(1024) k_bin_scan(CpvkBinArgs a) {
    __shared__ cpvk_u32 warpSums[32];
    __shared__ cpvk_u32 maxShared;
    const cpvk_u32 tiles = a.tilesX * a.tilesY;
    // thread t owns the contiguous run [t * per, (t + 1) * per): one serial pass for its sum, one block-wide scan of
    // the 1024 sums, one serial pass to write the offsets (three barriers in total, whatever the tile count)
    const cpvk_u32 per = (tiles + blockDim.x - 1) / blockDim.x;
    const cpvk_u32 begin = min(threadIdx.x * per, tiles), end = min(begin + per, tiles);
    if (threadIdx.x == 0) maxShared = 0;
    cpvk_u32 sum = 0, localMax = 0;
    // runs start at a multiple of `per`: when that is a multiple of four the counts come as 16-byte vectors (a 4K frame is
    // two loads per thread instead of a chain of eight)
    const bool vec = (per & 3u) == 0u && (reinterpret_cast<cpvk_u64>(a.counts) & 15u) == 0u;
    for (cpvk_u32 i = begin; i < end;) {
        if (vec && i + 4 <= end) {
            const uint4 v = *reinterpret_cast<const uint4*>(a.counts + i);
            sum += (v.x + v.y) + (v.z + v.w); localMax = max(max(localMax, max(v.x, v.y)), max(v.z, v.w)); i += 4;
        } else { const cpvk_u32 v = a.counts[i]; sum += v; localMax = max(localMax, v); i++; }
    }
    cpvk_u32 x = sum;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const cpvk_u32 y = __shfl_up_sync(0xFFFFFFFFu, x, d); if ((threadIdx.x & 31) >= d) x += y; }
    if ((threadIdx.x & 31) == 31) warpSums[threadIdx.x >> 5] = x;
    localMax = __reduce_max_sync(0xFFFFFFFFu, localMax);
    __syncthreads();
    if (threadIdx.x < 32) {
        cpvk_u32 w = warpSums[threadIdx.x];
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const cpvk_u32 y = __shfl_up_sync(0xFFFFFFFFu, w, d); if (threadIdx.x >= d) w += y; }
        warpSums[threadIdx.x] = w;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(&maxShared, localMax);
    __syncthreads();
    cpvk_u32 run = ((threadIdx.x >> 5) ? warpSums[(threadIdx.x >> 5) - 1] : 0u) + x - sum; // exclusive prefix of this thread's run
    for (cpvk_u32 i = begin; i < end;) {
        if (vec && i + 4 <= end && ((reinterpret_cast<cpvk_u64>(a.offsets) | reinterpret_cast<cpvk_u64>(a.cursors)) & 15u) == 0u) {
            const uint4 v = *reinterpret_cast<const uint4*>(a.counts + i);
            const uint4 o = make_uint4(run, run + v.x, run + v.x + v.y, run + v.x + v.y + v.z);
            *reinterpret_cast<uint4*>(a.offsets + i) = o; *reinterpret_cast<uint4*>(a.cursors + i) = o;
            run = o.w + v.w; i += 4;
        } else { a.offsets[i] = run; a.cursors[i] = run; run += a.counts[i]; i++; }
    }
    if (threadIdx.x == 0) {
        const cpvk_u32 total = warpSums[31], longest = maxShared;
        a.offsets[tiles] = total; a.meta[0] = total; a.meta[1] = longest;
        const bool fits = total <= a.planCapacity && (longest <= CPVK_ORDER_MAX || longest <= a.planSortCap) && (a.meta[2] == 0 || a.planLargeCounted != 0);
        a.meta[3] = fits ? 0u : 1u;
        if (a.metaHost) { // the host reads the answers from here after an event / stream sync: no copy in the stream
            a.metaHost[0] = total; a.metaHost[1] = longest; a.metaHost[2] = a.meta[2]; a.metaHost[3] = fits ? 0u : 1u;
            __threadfence_system();
        }
    }
}
// Per-tile ascending sort. Lists that fit the dynamic shared buffer use a bitonic network; longer ones fall back
// to an in-place stable LSD split sort through `scratch` (rare: > capacity primitives over one 32x32 tile).
__global__ void __launch_bounds__(256) k_bin_sort(CpvkBinArgs a, cpvk_u32 capacity) {
    extern __shared__ cpvk_u32 sKeys[];
    const cpvk_u32 t = blockIdx.x;
    if (a.meta[3] != 0 || a.meta[1] <= CPVK_ORDER_MAX) return; // plan mismatch, or every list is short enough to be ordered inside k_raster
    const cpvk_u32 begin = a.offsets[t], n = a.offsets[t + 1] - begin;
    if (n < 2) return;
    cpvk_u32* list = a.lists + begin;
    if (n <= capacity) {
        cpvk_u32 m = 1; while (m < n) m <<= 1;
        for (cpvk_u32 i = threadIdx.x; i < m; i += blockDim.x) sKeys[i] = i < n ? list[i] : 0xFFFFFFFFu;
        __syncthreads();
        for (cpvk_u32 k = 2; k <= m; k <<= 1)
            for (cpvk_u32 j = k >> 1; j > 0; j >>= 1) {
                for (cpvk_u32 i = threadIdx.x; i < m; i += blockDim.x) {
                    const cpvk_u32 l = i ^ j;
                    if (l > i) {
                        const cpvk_u32 x = sKeys[i], y = sKeys[l];
                        const bool up = (i & k) == 0;
                        if ((x > y) == up) { sKeys[i] = y; sKeys[l] = x; }
                    }
                }
                __syncthreads();
            }
        for (cpvk_u32 i = threadIdx.x; i < n; i += blockDim.x) list[i] = sKeys[i];
        return;
    }
    // fallback: one-bit stable splits, bit 0 .. highest set bit of primCount, ping-ponging list <-> scratch
    __shared__ cpvk_u32 sWarp[8], sZeros, sBaseZ, sBaseO;
    cpvk_u32* src = list; cpvk_u32* dst = a.scratch + begin;
    int bitsNeeded = 0; while ((a.primCount >> bitsNeeded) != 0) bitsNeeded++;
    for (int bit = 0; bit < bitsNeeded; bit++) {
        if (threadIdx.x == 0) sZeros = 0;
        __syncthreads();
        cpvk_u32 z = 0;
        for (cpvk_u32 i = threadIdx.x; i < n; i += blockDim.x) z += ((src[i] >> bit) & 1u) ^ 1u;
        atomicAdd(&sZeros, z);
        __syncthreads();
        if (threadIdx.x == 0) { sBaseZ = 0; sBaseO = sZeros; }
        __syncthreads();
        for (cpvk_u32 base = 0; base < n; base += blockDim.x) {
            const cpvk_u32 i = base + threadIdx.x;
            const bool valid = i < n;
            const cpvk_u32 key = valid ? src[i] : 0u;
            const bool one = valid && ((key >> bit) & 1u);
            const bool zero = valid && !one;
            const cpvk_u32 zm = __ballot_sync(0xFFFFFFFFu, zero), om = __ballot_sync(0xFFFFFFFFu, one);
            const cpvk_u32 lt = (1u << (threadIdx.x & 31)) - 1u;
            const int w = threadIdx.x >> 5;
            if ((threadIdx.x & 31) == 0) sWarp[w] = (__popc(zm) << 16) | __popc(om);
            __syncthreads();
            cpvk_u32 zBefore = 0, oBefore = 0, zAll = 0, oAll = 0;
            for (int q = 0; q < 8; q++) { const cpvk_u32 v = sWarp[q]; if (q < w) { zBefore += v >> 16; oBefore += v & 0xFFFFu; } zAll += v >> 16; oAll += v & 0xFFFFu; }
            if (zero) dst[sBaseZ + zBefore + __popc(zm & lt)] = key;
            if (one) dst[sBaseO + oBefore + __popc(om & lt)] = key;
            __syncthreads();
            if (threadIdx.x == 0) { sBaseZ += zAll; sBaseO += oAll; }
            __syncthreads();
        }
        cpvk_u32* tmp = src; src = dst; dst = tmp;
    }
    if (src != list) { for (cpvk_u32 i = threadIdx.x; i < n; i += blockDim.x) list[i] = src[i]; }
}

// ---- clear: every texel gets the same packed bytes, so pack once per thread and store 16 bytes at a time ----
__global__ void __launch_bounds__(256) k_clear(CpvkDevAttachment img, CpvkClearArgs c) {
    __shared__ __align__(16) cpvk_u8 pattern[48]; // lcm(texel, 16) bytes of repeated texel (texel in {1,2,3,4,8,16}); 48 covers 3-byte texels
    const cpvk_u32 texel = cpvk_texel_size(img.format);
    if (threadIdx.x == 0) {
        __align__(16) cpvk_u8 one[16];
        if (c.isDepthStencil) cpvk_set_depth_stencil(img.format, one, c.depth, c.stencil);
        else if (cpvk_format_is_int(img.format)) cpvk_set_pixel_int(img.format, one, c.u);
        else cpvk_set_pixel_f32_dyn(img.format, one, c.f);
        for (cpvk_u32 i = 0; i < 48; i++) pattern[i] = one[i % texel];
    }
    __syncthreads();
    const cpvk_u32 rowBytes = texel * img.width;
    cpvk_u8* base = reinterpret_cast<cpvk_u8*>(img.address);
    // D32_SFLOAT_S8_UINT is the one format whose texel has bytes SetPixel never writes (5 of 8, GlslFunctions.cpp:898-914)
    const cpvk_u32 written = img.format == 130 ? 5u : texel;
    const bool vec = ((img.address | img.rowPitch | rowBytes) & 15) == 0 && (16 % texel) == 0 && written == texel;
    if (vec) {
        const uint4 v = *reinterpret_cast<const uint4*>(pattern);
        if (img.rowPitch == rowBytes) { // one contiguous block
            uint4* d = reinterpret_cast<uint4*>(base);
            const cpvk_u64 total = ((cpvk_u64)rowBytes * img.height) >> 4, stride = (cpvk_u64)gridDim.x * blockDim.x;
            for (cpvk_u64 i = (cpvk_u64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) d[i] = v;
        } else {
            const cpvk_u32 per = rowBytes >> 4;
            for (cpvk_u32 r = blockIdx.x; r < img.height; r += gridDim.x) {
                uint4* d = reinterpret_cast<uint4*>(base + (cpvk_u64)r * img.rowPitch);
                for (cpvk_u32 q = threadIdx.x; q < per; q += blockDim.x) d[q] = v;
            }
        }
    } else {
        const cpvk_u64 total = (cpvk_u64)img.width * img.height;
        for (cpvk_u64 i = (cpvk_u64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (cpvk_u64)gridDim.x * blockDim.x) {
            const cpvk_u64 r = i / img.width, x = i - r * img.width;
            cpvk_u8* d = base + r * img.rowPitch + x * texel;
            for (cpvk_u32 k = 0; k < written; k++) d[k] = pattern[k];
        }
    }
}

__global__ void __launch_bounds__(256) k_copy_rows(cpvk_u8* dst, cpvk_u32 dstPitch, const cpvk_u8* src, cpvk_u32 srcPitch, cpvk_u32 rowBytes, cpvk_u32 rows) {
    const cpvk_u64 align = (cpvk_u64)dst | (cpvk_u64)src | dstPitch | srcPitch | rowBytes;
    const cpvk_u64 stride = (cpvk_u64)gridDim.x * blockDim.x, start = (cpvk_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if ((align & 15) == 0) {
        if (dstPitch == rowBytes && srcPitch == rowBytes) { // one contiguous block
            const cpvk_u64 total = ((cpvk_u64)rowBytes * rows) >> 4;
            for (cpvk_u64 i = start; i < total; i += stride) reinterpret_cast<uint4*>(dst)[i] = __ldg(reinterpret_cast<const uint4*>(src) + i);
        } else {
            const cpvk_u32 per = rowBytes >> 4;
            for (cpvk_u32 r = blockIdx.x; r < rows; r += gridDim.x)
                for (cpvk_u32 q = threadIdx.x; q < per; q += blockDim.x)
                    reinterpret_cast<uint4*>(dst + (cpvk_u64)r * dstPitch)[q] = __ldg(reinterpret_cast<const uint4*>(src + (cpvk_u64)r * srcPitch) + q);
        }
    } else {
        const cpvk_u64 total = (cpvk_u64)rowBytes * rows;
        for (cpvk_u64 i = start; i < total; i += stride) { const cpvk_u64 r = i / rowBytes, q = i - r * rowBytes; dst[r * dstPitch + q] = src[r * srcPitch + q]; }
    }
}

// vkCmdBlitImage, one 2-D colour region: the reference samples the source as a 3-D image with lod 1 on a one-level
// chain (-> level 0, min filter), clamp-to-edge, then SetPixel (CommandBuffer.cpp:75-226). This kernel replays that
// arithmetic with the per-axis work factored out: everything SampleImageOfLevel (ImageSampler.cpp:461-579) derives from
// one coordinate (texel index / second tap / lerp weight) depends only on the destination column or only on the
// destination row, so a thread computes its column's terms once, the CTA computes its rows' terms once, and the inner
// loop is just taps + lerps + pack. The z axis always lands on slice 0 twice with weight 0; since both z planes are
// then the same bits, the reference's last lerp is lerp(v, v, 0) and is evaluated as exactly that.
struct CpvkBlitAxis { int c0, c1; float t; int dst; };
#ifndef CPVK_BLIT_ROWS
#define CPVK_BLIT_ROWS 32
#endif
#ifndef CPVK_BLIT_BATCH
#define CPVK_BLIT_BATCH 2 /* NEAREST from 8-bit sources: rows whose texels are loaded before the first of them is converted and stored */
#endif
__device__ __forceinline__ CpvkBlitAxis cpvk_blit_axis(int i, int dst0, int dst1, int src0, int src1, cpvk_u32 srcSize, cpvk_u32 filter) {
    CpvkBlitAxis r;
    r.dst = dst1 < dst0 ? i + dst1 : i + dst0;
    const float coordTexel = ((float)r.dst + 0.5f - (float)dst0) * ((float)(src1 - src0) / (float)(dst1 - dst0)) + (float)src0;
    const float coord = coordTexel / (float)srcSize;
    if (filter == 0) { // NEAREST: floor(u * size + shift), shift = 0
        r.c0 = r.c1 = cpvk_wrap((cpvk_i32)floorf(coord * (float)srcSize + 0.0f), (cpvk_i32)srcSize, 2); r.t = 0.0f;
    } else {
        const float sc = coord * (float)srcSize - 0.5f;
        const cpvk_i32 raw = (cpvk_i32)floorf(sc);
        r.c1 = cpvk_wrap(raw + 1, (cpvk_i32)srcSize, 2); r.c0 = cpvk_wrap(raw, (cpvk_i32)srcSize, 2);
        r.t = sc - floorf(sc);
    }
    return r;
}
// Texel codecs of the blit, chosen at compile time for the formats render targets and textures mostly have (K8 = R8G8B8A8 / B8G8R8A8
// UNORM, K16F = R16G16B16A16_SFLOAT); KDYN goes through the run-time format switch.
enum { KDYN = 0, K8 = 1, K16F = 2 };
template <int K> __device__ __forceinline__ void cpvk_blit_load(cpvk_u32 format, const cpvk_u8* p, float v[4], const float* lut) {
    if (K == K8) {
        const cpvk_u32 t = *reinterpret_cast<const cpvk_u32*>(p);
        float b0, b1, b2, b3;
#ifndef CPVK_BLIT_NO_LUT
        if (lut) { b0 = lut[t & 0xFFu]; b1 = lut[(t >> 8) & 0xFFu]; b2 = lut[(t >> 16) & 0xFFu]; b3 = lut[t >> 24]; } // LINEAR: 16 decodes per texel, the table is cheaper
        else
#endif
        { b0 = cpvk_unorm8(t & 0xFFu); b1 = cpvk_unorm8((t >> 8) & 0xFFu); b2 = cpvk_unorm8((t >> 16) & 0xFFu); b3 = cpvk_unorm8(t >> 24); }
        v[0] = format == 37 ? b0 : b2; v[1] = b1; v[2] = format == 37 ? b2 : b0; v[3] = b3;
    } else if (K == K16F) {
        cpvk_unpack_half4(*reinterpret_cast<const uint2*>(p), v);
    } else cpvk_get_pixel_f32_dyn(format, p, v, lut);
}
__device__ __forceinline__ cpvk_u32 cpvk_blit_pack8(cpvk_u32 format, const float v[4]) {
    const cpvk_u32 r = cpvk_float_to_unorm(v[0], 255.0f), g = cpvk_float_to_unorm(v[1], 255.0f), b = cpvk_float_to_unorm(v[2], 255.0f), a = cpvk_float_to_unorm(v[3], 255.0f);
    return format == 37 ? (r | (g << 8) | (b << 16) | (a << 24)) : (b | (g << 8) | (r << 16) | (a << 24));
}
// TPT destination columns per thread (4 for NEAREST: an aligned run of four destination texels leaves as one (K8) or two (K16F)
// 16-byte stores; 1 for LINEAR, whose four taps and twelve double lerps per texel want the registers), CPVK_BLIT_ROWS rows per CTA:
// the per-column terms are computed once per thread, the per-row terms once per CTA.
template <int KS, int KD, int FILTER, int TPT> __global__ void __launch_bounds__(256) k_blit(CpvkBlitArgs b) {
    __shared__ float lut[256]; // (float)k / 255.0f by the IEEE divide itself, for the run-time format path
    __shared__ CpvkBlitAxis rows[CPVK_BLIT_ROWS];
    if (KS == KDYN || (KS == K8 && FILTER == 1)) lut[threadIdx.x] = (float)threadIdx.x / 255.0f;
    const float* const lutK = (KS == KDYN || (KS == K8 && FILTER == 1)) ? lut : nullptr;
    const int dstW = abs(b.dstX1 - b.dstX0), dstH = abs(b.dstY1 - b.dstY0);
    const cpvk_u32 stexel = cpvk_texel_size(b.src.format), dtexel = cpvk_texel_size(b.dst.format);
    const cpvk_u64 spitch = (cpvk_u64)stexel * b.src.width; // the sampler addresses a level as tightly packed rows (Formats.cpp:583-587)
    const cpvk_u8* src = reinterpret_cast<const cpvk_u8*>(b.src.address);
    const CpvkFormat fi = cpvk_format(b.src.format);
    const cpvk_u32 comps = fi.type == CPVK_FT_DEPTH ? 1u : fi.comps;
    const int x4 = (int)(blockIdx.x * blockDim.x + threadIdx.x) * TPT;
    CpvkBlitAxis cx[TPT];
    #pragma unroll
    for (int k = 0; k < TPT; k++) cx[k] = cpvk_blit_axis(x4 + k < dstW ? x4 + k : 0, b.dstX0, b.dstX1, b.srcX0, b.srcX1, b.src.width, FILTER);
    // the z axis: one destination "slice" 0, source slices [0, 1) of a depth-1 image
    const CpvkBlitAxis cz = cpvk_blit_axis(0, 0, 1, 0, 1, 1u, FILTER);
    // the four columns form one aligned in-range run of the destination? (then vector stores)
    const bool run = TPT == 4 && x4 + 3 < dstW && cx[0].dst >= 0 && (cpvk_u32)cx[TPT - 1].dst < b.dst.width && cx[TPT - 1].dst == cx[0].dst + 3 &&
                     (KD == K8 || KD == K16F) && ((b.dst.address + (cpvk_u64)cx[0].dst * dtexel) & 15u) == 0u && (b.dst.rowPitch & 15u) == 0u;
    const bool srcRun = TPT == 4 && KS == K8 && x4 + 3 < dstW && cx[1 % TPT].c0 == cx[0].c0 + 1 && cx[2 % TPT].c0 == cx[0].c0 + 2 && cx[3 % TPT].c0 == cx[0].c0 + 3 &&
                        ((b.src.address + (cpvk_u64)(cpvk_u32)cx[0].c0 * 4u) & 15u) == 0u && (spitch & 15u) == 0u;
    for (int rowBase = (int)blockIdx.y * CPVK_BLIT_ROWS; rowBase < dstH; rowBase += (int)gridDim.y * CPVK_BLIT_ROWS) {
        __syncthreads();
        if (threadIdx.x < CPVK_BLIT_ROWS && rowBase + (int)threadIdx.x < dstH)
            rows[threadIdx.x] = cpvk_blit_axis(rowBase + (int)threadIdx.x, b.dstY0, b.dstY1, b.srcY0, b.srcY1, b.src.height, FILTER);
        __syncthreads();
        if (x4 >= dstW) continue;
        const int nRows = min(CPVK_BLIT_ROWS, dstH - rowBase);
        auto store = [&](const CpvkBlitAxis& cy, float (&value)[TPT][4]) {
            if (cy.dst < 0 || (cpvk_u32)cy.dst >= b.dst.height) return;
            cpvk_u8* drow = reinterpret_cast<cpvk_u8*>(b.dst.address) + (cpvk_u64)cy.dst * b.dst.rowPitch;
            if (TPT == 4 && run && KD == K8) {
                *reinterpret_cast<uint4*>(drow + (cpvk_u64)cx[0].dst * 4u) = make_uint4(cpvk_blit_pack8(b.dst.format, value[0]), cpvk_blit_pack8(b.dst.format, value[1 % TPT]),
                                                                                       cpvk_blit_pack8(b.dst.format, value[2 % TPT]), cpvk_blit_pack8(b.dst.format, value[3 % TPT]));
            } else if (TPT == 4 && run && KD == K16F) {
                const uint2 h0 = cpvk_pack_half4(value[0]), h1 = cpvk_pack_half4(value[1 % TPT]), h2 = cpvk_pack_half4(value[2 % TPT]), h3 = cpvk_pack_half4(value[3 % TPT]);
                uint4* d = reinterpret_cast<uint4*>(drow + (cpvk_u64)cx[0].dst * 8u);
                d[0] = make_uint4(h0.x, h0.y, h1.x, h1.y); d[1] = make_uint4(h2.x, h2.y, h3.x, h3.y);
            } else {
                #pragma unroll
                for (int k = 0; k < TPT; k++) {
                    if (x4 + k >= dstW || cx[k].dst < 0 || (cpvk_u32)cx[k].dst >= b.dst.width) continue;
                    cpvk_set_pixel_f32_dyn(b.dst.format, drow + (cpvk_u64)cx[k].dst * dtexel, value[k]);
                }
            }
        };
        if (FILTER == 0 && KS == K8 && CPVK_BLIT_BATCH > 1) {
            // NEAREST from an 8-bit source is pure data movement: the texels of CPVK_BLIT_BATCH rows are requested before the first
            // one is converted, so that a thread has that many rows of loads in flight instead of one (read-only loads: the
            // stores of one row do not hold back the loads of the next)
            #pragma unroll 1
            for (int r = 0; r < nRows; r += CPVK_BLIT_BATCH) {
                cpvk_u32 raw[CPVK_BLIT_BATCH][TPT];
                #pragma unroll
                for (int j = 0; j < CPVK_BLIT_BATCH; j++) {
                    const cpvk_u8* r0 = src + (cpvk_u64)(cpvk_u32)rows[min(r + j, nRows - 1)].c0 * spitch;
                    if (TPT == 4 && srcRun) { // the four source texels are one aligned 16-byte run (an unscaled blit): one load
                        const uint4 q = __ldg(reinterpret_cast<const uint4*>(r0 + (cpvk_u64)(cpvk_u32)cx[0].c0 * 4u));
                        raw[j][0] = q.x; raw[j][1 % TPT] = q.y; raw[j][2 % TPT] = q.z; raw[j][3 % TPT] = q.w;
                        continue;
                    }
                    #pragma unroll
                    for (int k = 0; k < TPT; k++) raw[j][k] = x4 + k < dstW ? __ldg(reinterpret_cast<const cpvk_u32*>(r0 + (cpvk_u64)(cpvk_u32)cx[k].c0 * 4u)) : 0u;
                }
                #pragma unroll
                for (int j = 0; j < CPVK_BLIT_BATCH; j++) {
                    if (r + j >= nRows) break;
                    float value[TPT][4];
                    #pragma unroll
                    for (int k = 0; k < TPT; k++) {
                        const cpvk_u32 t = raw[j][k];
                        const float b0 = cpvk_unorm8(t & 0xFFu), b1 = cpvk_unorm8((t >> 8) & 0xFFu), b2 = cpvk_unorm8((t >> 16) & 0xFFu), b3 = cpvk_unorm8(t >> 24);
                        value[k][0] = b.src.format == 37 ? b0 : b2; value[k][1] = b1; value[k][2] = b.src.format == 37 ? b2 : b0; value[k][3] = b3;
                    }
                    store(rows[r + j], value);
                }
            }
            continue;
        }
        #pragma unroll 1
        for (int r = 0; r < nRows; r++) {
            const CpvkBlitAxis cy = rows[r];
            const cpvk_u8* r0 = src + (cpvk_u64)(cpvk_u32)cy.c0 * spitch;
            const cpvk_u8* r1 = src + (cpvk_u64)(cpvk_u32)cy.c1 * spitch;
            float value[TPT][4];
            #pragma unroll
            for (int k = 0; k < TPT; k++) {
                if (x4 + k >= dstW) continue;
                float* v = value[k];
                if (FILTER == 0) {
                    cpvk_blit_load<KS>(b.src.format, r0 + (cpvk_u64)(cpvk_u32)cx[k].c0 * stexel, v, lutK);
                } else {
                    CpvkVec4 i0j0, i1j0, i0j1, i1j1;
                    cpvk_blit_load<KS>(b.src.format, r0 + (cpvk_u64)(cpvk_u32)cx[k].c0 * stexel, i0j0.v, lutK);
                    cpvk_blit_load<KS>(b.src.format, r0 + (cpvk_u64)(cpvk_u32)cx[k].c1 * stexel, i1j0.v, lutK);
                    cpvk_blit_load<KS>(b.src.format, r1 + (cpvk_u64)(cpvk_u32)cx[k].c0 * stexel, i0j1.v, lutK);
                    cpvk_blit_load<KS>(b.src.format, r1 + (cpvk_u64)(cpvk_u32)cx[k].c1 * stexel, i1j1.v, lutK);
                    const CpvkVec4 ij0 = cpvk_lerp(i0j0, i1j0, cx[k].t), ij1 = cpvk_lerp(i0j1, i1j1, cx[k].t);
                    const CpvkVec4 plane = cpvk_lerp(ij0, ij1, cy.t);
                    const CpvkVec4 out = cpvk_lerp_same(plane, cz.t); // the two z planes are the same slice
                    v[0] = out.v[0]; v[1] = out.v[1]; v[2] = out.v[2]; v[3] = out.v[3];
                }
                if (comps < 2) v[1] = 0.0f; // SampleImage (ImageSampler.cpp:581-673): absent channels read 0, 0, 1
                if (comps < 3) v[2] = 0.0f;
                if (comps < 4) v[3] = 1.0f;
            }
            store(cy, value);
        }
    }
}

// Diagnostics: cpvk_div_shared (cpvk_device.cuh) on caller-supplied operands, three numerators per denominator, next to the
// plain operator — tests/test_parity_gpu.py compares both with IEEE division computed on the host over every exponent range.
__global__ void __launch_bounds__(256) k_selftest_div(const float* a, const float* b, cpvk_u32 n, float* shared, float* plain) {
    const cpvk_u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v[3] = {a[3 * i], a[3 * i + 1], a[3 * i + 2]};
    const float d = b[i];
    #pragma unroll
    for (int k = 0; k < 3; k++) plain[3 * i + k] = v[k] / d;
    cpvk_div_shared(v, d);
    #pragma unroll
    for (int k = 0; k < 3; k++) shared[3 * i + k] = v[k];
}

// ---- host-callable launchers (cpvk_abi.cpp is plain C++) ----
// cpvk_cuda_peer_barrier (include/cpvk_cuda.h): lane i signals participant i, then waits for participant i's signal.
struct CpvkPeerFlags { cpvk_u32* p[16]; };
__global__ void __launch_bounds__(32) k_peer_barrier(CpvkPeerFlags f, cpvk_u32 count, cpvk_u32 self, cpvk_u32 sequence, const cpvk_u32* verdict) {
    const cpvk_u32 i = threadIdx.x;
    if (i >= count || i == self) return;
    if (verdict && verdict[3] != 0) return; // the draw in front of this barrier was a no-op (plan mismatch): the host replays it, then the barrier
    // the kernels before this one in the stream are complete, their stores (peer memory included) performed; the release
    // makes the flag the last thing a peer can see of them
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(f.p[i] + self), "r"(sequence) : "memory");
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        cpvk_u32 v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f.p[self] + i) : "memory");
        if ((cpvk_i32)(v - sequence) >= 0) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > 10000000000ull) __trap(); // a participant never arrived: fail loudly rather than hang the GPU
        __nanosleep(40);
    }
}

static inline unsigned cpvk_grid(unsigned long long n, unsigned block) { return (unsigned)((n + block - 1) / block); }

extern "C" {
cudaError_t cpvk_launch_setup(const CpvkSetupArgs* a, cudaStream_t s) {
    if (a->primCount == 0) return cudaSuccess;
    k_setup<<<cpvk_grid(a->primCount, 256), 256, 0, s>>>(*a);
    return cudaGetLastError();
}
cudaError_t cpvk_launch_bin(const CpvkBinArgs* a, int pass, int small, int large, cudaStream_t s) {
    if (a->primCount == 0) return cudaSuccess;
    if (small) k_bin<<<cpvk_grid(a->primCount, 256), 256, 0, s>>>(*a, pass);
    if (large) k_bin_large<<<592, 256, 0, s>>>(*a, pass); // 148 SMs x 4 resident CTAs; loops over the deferred list
    return cudaGetLastError();
}
cudaError_t cpvk_launch_index_range(unsigned long long indexBuffer, unsigned indexStride, unsigned first, unsigned count, cpvk_u32* range, cudaStream_t s) {
    if (!count) return cudaSuccess;
    unsigned grid = cpvk_grid(count, 256 * 8);
    if (grid > 148 * 8) grid = 148 * 8;
    k_index_range<<<grid, 256, 0, s>>>(indexBuffer, indexStride, first, count, range);
    return cudaGetLastError();
}
cudaError_t cpvk_launch_bin_scan(const CpvkBinArgs* a, cudaStream_t s) {
    k_bin_scan<<<1, 1024, 0, s>>>(*a);
    return cudaGetLastError();
}
cudaError_t cpvk_launch_bin_sort(const CpvkBinArgs* a, unsigned capacity, cudaStream_t s) {
    if (capacity * 4 > 48 * 1024) cudaFuncSetAttribute(k_bin_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    k_bin_sort<<<a->tilesX * a->tilesY, 256, capacity * 4, s>>>(*a, capacity);
    return cudaGetLastError();
}
cudaError_t cpvk_launch_clear(const CpvkDevAttachment* img, const CpvkClearArgs* c, cudaStream_t s) {
    const unsigned long long texels = (unsigned long long)img->width * img->height;
    unsigned grid = cpvk_grid(texels, 256 * 4);
    if (grid > 148 * 16) grid = 148 * 16;
    if (grid == 0) return cudaSuccess;
    k_clear<<<grid, 256, 0, s>>>(*img, *c);
    return cudaGetLastError();
}
cudaError_t cpvk_launch_copy_rows(unsigned long long dst, unsigned dstPitch, unsigned long long src, unsigned srcPitch, unsigned rowBytes, unsigned rows, cudaStream_t s) {
    if (!rowBytes || !rows) return cudaSuccess;
    unsigned grid = cpvk_grid((unsigned long long)rowBytes * rows / 16 + 1, 256);
    if (grid > 148 * 16) grid = 148 * 16;
    k_copy_rows<<<grid, 256, 0, s>>>((cpvk_u8*)dst, dstPitch, (const cpvk_u8*)src, srcPitch, rowBytes, rows);
    return cudaGetLastError();
}
cudaError_t cpvk_launch_peer_barrier(const unsigned long long* flagArrays, unsigned count, unsigned self, unsigned sequence, const cpvk_u32* verdict, cudaStream_t s) {
    CpvkPeerFlags f{};
    for (unsigned i = 0; i < count && i < 16; i++) f.p[i] = reinterpret_cast<cpvk_u32*>(flagArrays[i]);
    k_peer_barrier<<<1, 32, 0, s>>>(f, count, self, sequence, verdict);
    return cudaGetLastError();
}
cudaError_t cpvk_launch_selftest_div(const float* a, const float* b, unsigned n, float* shared, float* plain, cudaStream_t s) {
    if (!n) return cudaSuccess;
    k_selftest_div<<<cpvk_grid(n, 256), 256, 0, s>>>(a, b, n, shared, plain);
    return cudaGetLastError();
}
cudaError_t cpvk_launch_blit(const CpvkBlitArgs* b, cudaStream_t s) {
    const unsigned long long total = (unsigned long long)abs(b->dstX1 - b->dstX0) * (unsigned long long)abs(b->dstY1 - b->dstY0);
    if (!total) return cudaSuccess;
    const unsigned w = (unsigned)abs(b->dstX1 - b->dstX0), h = (unsigned)abs(b->dstY1 - b->dstY0);
    const int lin = b->filter != 0;
    dim3 grid(cpvk_grid(cpvk_grid(w, lin ? 1 : 4), 256), cpvk_grid(h, CPVK_BLIT_ROWS));
    if (grid.y > 65535u) grid.y = 65535u; // the kernel strides over row blocks
    auto kind = [](unsigned f) { return (f == 37 || f == 44) ? K8 : (f == 97 ? K16F : KDYN); };
    const int ks = kind(b->src.format), kd = kind(b->dst.format);
    #define CPVK_BLIT_CASE(S, D) if (ks == S && kd == D) { if (lin) k_blit<S, D, 1, 1><<<grid, 256, 0, s>>>(*b); else k_blit<S, D, 0, 4><<<grid, 256, 0, s>>>(*b); return cudaGetLastError(); }
    CPVK_BLIT_CASE(K8, K8) CPVK_BLIT_CASE(K8, K16F) CPVK_BLIT_CASE(K16F, K8) CPVK_BLIT_CASE(K16F, K16F)
    #undef CPVK_BLIT_CASE
    if (lin) k_blit<KDYN, KDYN, 1, 1><<<grid, 256, 0, s>>>(*b); else k_blit<KDYN, KDYN, 0, 4><<<grid, 256, 0, s>>>(*b);
    return cudaGetLastError();
}
}
