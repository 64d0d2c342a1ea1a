// apply.cu -- label application: the hand-off after inference.
//   Tracking::DynamicDetectionWithCRF   src/Tracking.cc:1945-1955
//     for (i < N) if (res_label[i] == 0) { maps.erase(fid); pMP->SetBadFlag(); mvpMapPoints[fid] = NULL; }
// The pointer surgery stays on the host; what the device produces is the list the loop walks: the feature ids
// (or local point indices) of the points labelled moving, in point order, per problem -- and the complementary
// survivor list that the second PoseOptimization consumes (Tracking.cc:1002).  Both are a STABLE partition of the
// batch's MAP labels, so a host loop over dyn_list visits exactly the elements the reference loop visits, in the
// same order.
//
// Three small HBM streams over the batch: tile counts -> exclusive scan of the tile counts -> scatter.
#include "engine.cuh"

namespace lccrf {

namespace {

constexpr int kPartThreads = 256;
constexpr int kPartItems = 8;  // points per thread
constexpr int kPartTile = kPartThreads * kPartItems;

__device__ __forceinline__ int block_exclusive_scan(int v, int *s_warp, int &total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[wid] = x;
    __syncthreads();
    int wsum = lane < kPartThreads / 32 ? s_warp[lane] : 0;
    int wx = wsum;
#pragma unroll
    for (int o = 1; o < kPartThreads / 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, wx, o);
        if (lane >= o) wx += y;
    }
    total = __shfl_sync(0xffffffffu, wx, kPartThreads / 32 - 1);
    const int wbase = __shfl_sync(0xffffffffu, wx - wsum, wid);
    __syncthreads();
    return wbase + x - v;
}

// moving points (label 0) of every tile
__global__ void __launch_bounds__(kPartThreads)
k_part_count(const short *__restrict__ map, int NT, int *__restrict__ tile_dyn) {
    __shared__ int s_warp[kPartThreads / 32];
    const int base = blockIdx.x * kPartTile + threadIdx.x * kPartItems;
    int c = 0;
#pragma unroll
    for (int q = 0; q < kPartItems; q++)
        if (base + q < NT) c += map[base + q] == 0;
    int total;
    block_exclusive_scan(c, s_warp, total);
    if (threadIdx.x == 0) tile_dyn[blockIdx.x] = total;
}

// exclusive scan of the tile counts (one CTA; the batch has NT / 2048 tiles), total at tile_off[tiles]
__global__ void __launch_bounds__(kPartThreads)
k_part_scan(const int *__restrict__ tile_dyn, int tiles, int *__restrict__ tile_off) {
    __shared__ int s_warp[kPartThreads / 32];
    int carry = 0;
    for (int t0 = 0; t0 < tiles; t0 += kPartThreads) {
        const int t = t0 + threadIdx.x;
        const int v = t < tiles ? tile_dyn[t] : 0;
        int total;
        const int ex = block_exclusive_scan(v, s_warp, total);
        if (t < tiles) tile_off[t] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) tile_off[tiles] = carry;
}

// stable scatter; the thread that owns the first point of a problem also records the problem's list starts
__global__ void __launch_bounds__(kPartThreads)
k_part_scatter(const short *__restrict__ map, int NT, const int *__restrict__ tile_off, int tiles,
               const int *__restrict__ prob_ptr, int B, const int *__restrict__ fid, int *__restrict__ dyn_ptr,
               int *__restrict__ dyn_list, int *__restrict__ stat_ptr, int *__restrict__ stat_list) {
    __shared__ int s_warp[kPartThreads / 32];
    const int base = blockIdx.x * kPartTile + threadIdx.x * kPartItems;
    short m[kPartItems];
    int c = 0;
#pragma unroll
    for (int q = 0; q < kPartItems; q++) {
        m[q] = base + q < NT ? map[base + q] : (short)1;
        c += m[q] == 0;
    }
    int total;
    int dyn_before = tile_off[blockIdx.x] + block_exclusive_scan(c, s_warp, total);
    if (base < NT) {
        int b = find_segment(prob_ptr, B + 1, base);  // last b with prob_ptr[b] <= base
        int pstart = __ldg(prob_ptr + b), pend = __ldg(prob_ptr + b + 1);
#pragma unroll
        for (int q = 0; q < kPartItems; q++) {
            const int i = base + q;
            if (i >= NT) break;
            while (i >= pend) {  // crossed into the next non-empty problem
                b++;
                pstart = pend;
                pend = __ldg(prob_ptr + b + 1);
            }
            if (i == pstart) {  // list starts of problem b and of the empty problems in front of it
                for (int bb = b; bb >= 0 && __ldg(prob_ptr + bb) == i; bb--) {
                    dyn_ptr[bb] = dyn_before;
                    stat_ptr[bb] = i - dyn_before;
                }
            }
            const int v = fid ? __ldg(fid + i) : i - pstart;
            if (m[q] == 0) dyn_list[dyn_before++] = v;
            else stat_list[i - dyn_before] = v;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // list ends, and the starts of trailing empty problems
        const int all = tile_off[tiles];
        for (int bb = B; bb >= 0 && __ldg(prob_ptr + bb) == NT; bb--) {
            dyn_ptr[bb] = all;
            stat_ptr[bb] = NT - all;
        }
    }
}

}  // namespace

size_t label_partition_scratch_bytes(int NT) { return ((size_t)cdiv(NT > 0 ? NT : 1, kPartTile) + 1) * 2 * sizeof(int); }

// map [NT] (0 = moving) -> dyn_list / stat_list [NT] and dyn_ptr / stat_ptr [B+1]; fid optional [NT]; all device pointers
int label_partition(Ctx *ctx, const short *map, int NT, const int *prob_ptr, int B, const int *fid, int *scratch,
                    int *dyn_ptr, int *dyn_list, int *stat_ptr, int *stat_list) {
    const int tiles = cdiv(NT > 0 ? NT : 1, kPartTile);
    int *tile_dyn = scratch, *tile_off = scratch + tiles;
    { LCCRF_KERNEL(ctx, "k_part_count"); k_part_count<<<tiles, kPartThreads, 0, ctx->stream>>>(map, NT, tile_dyn); }
    { LCCRF_KERNEL(ctx, "k_part_scan"); k_part_scan<<<1, kPartThreads, 0, ctx->stream>>>(tile_dyn, tiles, tile_off); }
    { LCCRF_KERNEL(ctx, "k_part_scatter"); k_part_scatter<<<tiles, kPartThreads, 0, ctx->stream>>>(
        map, NT, tile_off, tiles, prob_ptr, B, fid, dyn_ptr, dyn_list, stat_ptr, stat_list); }
    LCCRF_CUDA(cudaGetLastError());
    return LCCRF_OK;
}

}  // namespace lccrf
