// topk.cuh -- warp-cooperative exact top-k selection shared by the kNN kernels.
//
// One warp owns one query row.  The current best-k set is kept UNSORTED, one entry per lane (two for
// k > 32), together with the warp-uniform worst value `wmin`.  Candidates arrive one per lane per step in
// increasing index order, so a candidate whose value ties the worst kept value never enters (it has the
// higher index) and the only tie rule needed on eviction is "among equal worst values drop the largest
// index".  The set is sorted once at the end by (value desc, index asc) with a bitonic network on packed
// 64-bit keys.  This realises the reference ranking `pairwise_distance.topk(k)` (PointDA/model_utils.py:15)
// with the lowest-index tie rule fixed by BASELINE.json:north_star.
#pragma once
#include "common.cuh"

namespace mlsp {

// 64-bit ranking key: ascending key order == (value descending, index ascending); dead entries last.
__device__ __forceinline__ unsigned long long rank_key(float val, int idx, bool live)
{
    const float c = __fadd_rn(val, 0.0f);  // -0 -> +0 so equal values give equal keys
    const uint32_t hi = live ? ~f32_orderable(c) : 0xffffffffu;
    return ((unsigned long long)hi << 32) | (uint32_t)idx;
}

// Bitonic sort of 32*KSLOTS keys held as key[s] on lane l <-> element e = s*32 + l, ascending.
template <int KSLOTS>
__device__ __forceinline__ void warp_sort_u64(unsigned long long (&key)[KSLOTS])
{
    const int lane = lane_id();
#pragma unroll
    for (int size = 2; size <= 32 * KSLOTS; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 32) {
                const int ds = stride / 32;  // partner slot distance; direction depends on the slot only
#pragma unroll
                for (int s = 0; s < KSLOTS; ++s) {
                    if ((s & ds) == 0) {
                        const bool up = ((s * 32) & size) == 0;
                        const unsigned long long a = key[s], b = key[s | ds];
                        const bool a_small = a < b;
                        key[s] = (a_small == up) ? a : b;
                        key[s | ds] = (a_small == up) ? b : a;
                    }
                }
            } else {
#pragma unroll
                for (int s = 0; s < KSLOTS; ++s) {
                    const unsigned long long other = __shfl_xor_sync(MLSP_FULL, key[s], stride);
                    const int e = s * 32 + lane;
                    const bool up = (e & size) == 0;
                    const bool lower = (lane & stride) == 0;
                    const bool take_min = (lower == up);
                    const bool other_smaller = other < key[s];
                    key[s] = (take_min == other_smaller) ? other : key[s];
                }
            }
        }
    }
}

template <int KSLOTS>
struct TopK {
    float v[KSLOTS];
    int j[KSLOTS];
    float wmin;  // warp-uniform: worst value currently kept (-inf while the set is not full)
    int wj;      // warp-uniform: index of the entry to evict next (largest index among value == wmin)

    // Live slots start as (-inf, unique huge index) so the (wmin, wj) pair always names exactly one
    // entry; slots with element index s*32+lane >= k are dead: (+inf, -1), never the minimum.
    __device__ __forceinline__ void init(int k)
    {
#pragma unroll
        for (int s = 0; s < KSLOTS; ++s) {
            const int e = s * 32 + lane_id();
            v[s] = e < k ? -INFINITY : INFINITY;
            j[s] = e < k ? 0x7fffffff - e : -1;
        }
        wmin = -INFINITY;
        wj = 0x7fffffff;
    }

    __device__ __forceinline__ bool slot_live(int s, int k) const { return s * 32 + lane_id() < k; }

    // Offer one candidate per lane (pd, idx); `valid` masks lanes without a candidate.
    __device__ __forceinline__ void offer(float pd, int idx, bool valid)
    {
        bool pass = valid && (pd > wmin);
        unsigned m = __ballot_sync(MLSP_FULL, pass);
        while (m) {
            const int src = __ffs(m) - 1;
            const float cv = __shfl_sync(MLSP_FULL, pd, src);
            const int cj = __shfl_sync(MLSP_FULL, idx, src);
#pragma unroll
            for (int s = 0; s < KSLOTS; ++s)
                if (v[s] == wmin && j[s] == wj) {
                    v[s] = cv;
                    j[s] = cj;
                }
            float lo = v[0];
#pragma unroll
            for (int s = 1; s < KSLOTS; ++s) lo = fminf(lo, v[s]);
            wmin = warp_min_f32(lo);
            int cand = -1;
#pragma unroll
            for (int s = 0; s < KSLOTS; ++s)
                if (v[s] == wmin) cand = max(cand, j[s]);
            wj = __reduce_max_sync(MLSP_FULL, cand);
            pass = pass && (lane_id() != src) && (pd > wmin);
            m = __ballot_sync(MLSP_FULL, pass);
        }
    }

    // Sort the kept set: best first (value descending, index ascending).  Afterwards element
    // e = s*32+lane of the ranking is in j[s] (only e < k is meaningful).
    __device__ __forceinline__ void finish(int k)
    {
        unsigned long long key[KSLOTS];
#pragma unroll
        for (int s = 0; s < KSLOTS; ++s) key[s] = rank_key(v[s], j[s], slot_live(s, k));
        warp_sort_u64<KSLOTS>(key);
#pragma unroll
        for (int s = 0; s < KSLOTS; ++s) j[s] = (int)(uint32_t)(key[s] & 0xffffffffull);
    }
};

}  // namespace mlsp
