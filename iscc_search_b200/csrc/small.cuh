// small.cuh - the small-batch (HBM-bound) search path: ONE persistent cooperative kernel per search.
//
// A request of iscc-search carries one asset, i.e. one query per unit-type index
// (/root/reference/iscc_search/indexes/usearch/index.py:786-806, 2024-2045): the store is streamed once for one (or a
// handful of) queries and the scan is bound by HBM bandwidth, not by the POPC pipe. For that regime the multi-launch
// plan of the batch path (gather, init, sample, sample_tau, one scan launch per compared length, select: 7-9 launches
// and their gaps) costs 15-30 % of the request. k_scan_small does the whole search in one launch:
//
//   phase 0  threshold bootstrap: every CTA histograms a few units spread evenly over the store (no emission)
//            -> grid barrier -> tau = k-th rank of the sample (an upper bound of the true k-th rank)
//   phase 1  the scan: one CTA per SM, a producer warp feeds a ring of shared-memory stages with TMA bulk copies
//            (cp.async.bulk + mbarrier complete_tx), 8 consumer warps score the rows of a stage against all (<= 8)
//            queries of ANY length in one pass, candidates within the running threshold are emitted exactly as in
//            k_scan; the producer warp also re-tightens the thresholds from the rank histograms
//   phase 2  grid barrier -> CTA q cuts query q at its exact k-th rank, breaks ties by key and writes the result
//
// Units: the block list is bucket ordered; a unit is up to 8 plane chunks of 4 KB (1024 rows of one 32-bit word
// plane): 1 block of a 5..8-word compare, 2 blocks of 3..4 words, 4 blocks of 2 words, 8 blocks of 1 word, so every
// stage moves the same number of bytes. The producer warps claim units from a global counter (one atomic per 64 KB unit,
// fetched one unit ahead), so no SM idles at the end of the scan; a stage with 0 blocks tells the consumers to stop.
#pragma once
#include <cooperative_groups.h>

#include "kernels.cuh"

namespace isx {

#define ISX_STAMP(i) do { if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 16 + (i)] = clock64(); } while (0)

constexpr int kSmallT = 8;                 // queries per launch
constexpr int kSmallRingBytes = 192 * 1024; // ring of S stages x CH plane chunks of 4 KB: (S, CH) = (5, 8) or (3, 16) - wider stages
                                            // mean fewer, larger TMA copies per byte (one per plane and stage)
constexpr int kSmallMaxChunks = 16;
constexpr int kSmallConsumers = 256;       // 8 consumer warps: thread t owns rows 4t..4t+3 of every block
constexpr int kSmallThreads = kSmallConsumers + 64;  // + the TMA producer warp + the threshold warp
constexpr int kSmallRanges = 8;            // length buckets per launch (ISCC has 4)
constexpr uint32_t kSmallMaxR = 1024;      // rank classes (385 for 64..256-bit codes)
constexpr uint32_t kSmallSortCap = 4096;   // survivors the fused select sorts in shared memory
constexpr uint32_t kChunkBytes = 4096;

// one per live 1024-row block, same order as ScanParams::blocks (built by upload_blocks)
struct BlockDesc {
    const uint32_t* base;  // plane 0 of the block's first row
    uint32_t cap;          // plane stride in rows
    uint32_t seg;          // segment id
    uint32_t row0;         // first row of the block inside its segment
    uint32_t seg_n;        // live rows of the segment
    uint32_t pad[2];
};
static_assert(sizeof(BlockDesc) == 32, "BlockDesc is read as two 16-byte vectors");

struct SmallRange {
    uint32_t unit0, n_units;   // units of this bucket: [unit0, unit0 + n_units)
    uint32_t block0, n_blocks; // its blocks in the block list
    uint32_t we;               // plane chunks per block = ceil(min(longest query, bucket length) / 4)
    uint32_t bpu;              // blocks per unit = CH / we
    uint32_t len_bytes;        // bucket length
    uint32_t pad;
};

struct SmallParams {
    const BlockDesc* bdesc;
    const SegDesc* segs;
    SmallRange ranges[kSmallRanges];
    uint32_t n_ranges, n_units;
    uint32_t T;
    uint32_t qwords[kSmallT][8];   // zero padded little-endian words (host queries), or
    const uint32_t* d_queries;     // device queries (caller order, q x 8 words), masked to qlen here
    uint32_t qlen[kSmallT];        // bytes
    uint32_t qsrc[kSmallT];        // index of the query in the caller's arrays (device queries + output row)
    // per-query state (clean between searches: tau = ~0, everything else 0)
    uint32_t* tau; uint32_t* hist; uint32_t* shist; uint32_t* cand_cnt; uint32_t* overflow;
    uint64_t* cand;
    uint32_t C, R, k, tau_init;
    const uint16_t* rank_tab;      // [33][257]
    const uint16_t* hmax_tab;      // [33][R]
    uint32_t sample_units;         // units per CTA in phase 0
    uint32_t sample_stride;        // sampled unit = (cta + j * grid) * stride
    // outputs [Q][k] (device or mapped pinned host memory)
    uint64_t* out_khi; uint64_t* out_klo; uint16_t* out_h; uint16_t* out_n; uint32_t* out_cnt; uint8_t* out_codes;
    uint32_t key_words;
    uint32_t* unit_counter;        // phase 1 units are claimed dynamically (0 between searches)
    unsigned long long* dbg;       // optional [gridDim][16] clock64 stamps (ISX_SMALL_DEBUG=1)
    uint32_t l2_evict_first;       // stream the rows with an evict-first L2 policy
    uint32_t static_units;         // 1: unit i of CTA c is c + i * gridDim (A/B against the dynamic claim)
    uint32_t* info;                // [T][4]: status (0 done, 1 overflow -> exact re-scan, 2 -> general select), candidates, d*, rows within d*
};

// ---- mbarrier / TMA bulk primitives (PTX ISA 8.x, sm_90+) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s_plain(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// rows are streamed once per search: evict-first keeps the histograms / candidate lists / thresholds resident in L2
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}

// shared-memory layout of k_scan_small (dynamic)
template <int S, int CH>
struct SmallSharedT {
    static_assert(S * CH * (int)kChunkBytes <= kSmallRingBytes, "ring too large");
    alignas(128) uint8_t ring[S][CH][kChunkBytes];                        // <= 160 KB; phase 2 reuses it as the sort buffer
    alignas(16) uint32_t qw[kSmallT][8];                                  // query words
    alignas(16) uint32_t qmask[kSmallRanges][kSmallT][8];                 // AND mask of every word for (bucket, query)
    uint32_t m[kSmallRanges][kSmallT];                                    // bytes compared for (bucket, query)
    uint32_t hcut[kSmallRanges][kSmallT];                                 // phase 0 counts only rows with h <= hcut (lower tail)
    alignas(8) uint64_t full[S], empty[S];
    uint32_t stage_range[S], stage_nblk[S];
    uint32_t stage_seg[S][CH], stage_row0[S][CH], stage_segn[S][CH];
    volatile uint32_t tau[kSmallT];                                       // current rank thresholds (threshold warp refreshes)
    volatile uint32_t hm[kSmallRanges][kSmallT];                          // largest Hamming distance within tau for (bucket, query)
    uint32_t dirty;                                                       // bit q: query q emitted since the last tighten
    volatile uint32_t stop;                                               // the producer issued its last unit: threshold warp exits
    uint32_t scan[16];
    uint32_t sel[8];
    // followed by: uint32_t s_hist[T][R] (phase 0), uint16_t hrow[n_ranges][T][R] is NOT kept: hmax comes from hmax_tab (L1)
};

// Batched form of tighten_tau for one warp: all bins <= tcur are fetched at once (R <= 1024: 32 per lane).
__device__ __forceinline__ uint32_t small_kth_rank(const uint32_t* hq, uint32_t tcur, uint32_t k, uint32_t lane) {
    // lane l owns bins [l*chunk, (l+1)*chunk): contiguous, so one shuffle scan over the lane sums finds the lane of the k-th row
    const uint32_t nb = tcur + 1;
    const uint32_t chunk = (nb + 31) / 32;
    uint32_t v[kSmallMaxR / 32];
    uint32_t sum = 0;
#pragma unroll
    for (uint32_t j = 0; j < kSmallMaxR / 32; j++) {
        const uint32_t r = lane * chunk + j;
        v[j] = (j < chunk && r < nb) ? __ldcg(&hq[r]) : 0u;
        sum += v[j];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    const unsigned hit = __ballot_sync(0xffffffffu, incl >= k);
    if (!hit) return 0xffffffffu;                 // fewer than k rows counted so far
    const int L = __ffs(hit) - 1;
    uint32_t res = 0xffffffffu;
    if ((int)lane == L) {
        uint32_t cum = incl - sum;
#pragma unroll
        for (uint32_t j = 0; j < kSmallMaxR / 32; j++) {
            cum += v[j];
            if (res == 0xffffffffu && cum >= k) res = lane * chunk + j;
        }
    }
    return __shfl_sync(0xffffffffu, res, L);
}

template <int WE, int S, int CH>
__device__ __forceinline__ void small_consume(const SmallParams& p, SmallSharedT<S, CH>& sh, uint32_t stage, uint32_t r, uint32_t nblk, bool sampling,
                                              uint32_t* s_hist, unsigned char* dirty_dummy, uint32_t& dirty_bits) {
    constexpr int BPU = CH / WE;   // blocks of this word count a stage holds; chunk (w, j) sits at ring[stage][w * BPU + j]
    const uint32_t tid = threadIdx.x;
    // all rows of the stage this thread owns (4 per block) go to registers first: the loads overlap
    uint4 a[BPU][WE];
#pragma unroll
    for (int j = 0; j < BPU; j++) {
#pragma unroll
        for (int w = 0; w < WE; w++) a[j][w] = *reinterpret_cast<const uint4*>(&sh.ring[stage][w * BPU + j][tid * 16]);
    }
#pragma unroll 1
    for (uint32_t q = 0; q < p.T; q++) {
        uint32_t qv[8], mk[8];
        {
            const uint4 v0 = *reinterpret_cast<const uint4*>(&sh.qw[q][0]);
            const uint4 m0 = *reinterpret_cast<const uint4*>(&sh.qmask[r][q][0]);
            qv[0] = v0.x; qv[1] = v0.y; qv[2] = v0.z; qv[3] = v0.w;
            mk[0] = m0.x; mk[1] = m0.y; mk[2] = m0.z; mk[3] = m0.w;
            if (WE > 4) {
                const uint4 v1 = *reinterpret_cast<const uint4*>(&sh.qw[q][4]);
                const uint4 m1 = *reinterpret_cast<const uint4*>(&sh.qmask[r][q][4]);
                qv[4] = v1.x; qv[5] = v1.y; qv[6] = v1.z; qv[7] = v1.w;
                mk[4] = m1.x; mk[5] = m1.y; mk[6] = m1.z; mk[7] = m1.w;
            }
        }
        const uint32_t hmax = sampling ? sh.hcut[r][q] : sh.hm[r][q];
        // distances of all blocks first (independent chains), ONE vote per (stage, query): rows within the threshold are rare
        uint32_t d[BPU][4];
        uint32_t dmin = 0xffffffffu;
#pragma unroll
        for (int j = 0; j < BPU; j++) {
#pragma unroll
            for (int rr = 0; rr < 4; rr++) {
                uint32_t x[WE];
#pragma unroll
                for (int w = 0; w < WE; w++) x[w] = (comp(a[j][w], rr) ^ qv[w]) & mk[w];
                d[j][rr] = pair_distance<WE>(x);
            }
            const uint32_t dj = min(min(d[j][0], d[j][1]), min(d[j][2], d[j][3]));
            dmin = min(dmin, j < (int)nblk ? dj : 0xffffffffu);   // blocks beyond nblk hold stale bytes of an earlier stage
        }
        if (!__any_sync(0xffffffffu, dmin <= hmax)) continue;
        const uint32_t m = sh.m[r][q];
#pragma unroll
        for (int j = 0; j < BPU; j++) {
            if (j >= (int)nblk) break;
            const uint32_t dj = min(min(d[j][0], d[j][1]), min(d[j][2], d[j][3]));
            if (!__any_sync(0xffffffffu, dj <= hmax)) continue;
            const uint32_t seg = sh.stage_seg[stage][j], row0 = sh.stage_row0[stage][j] + tid * 4, seg_n = sh.stage_segn[stage][j];
            if (sampling) {
                // only the lower tail matters for the k-th rank of the sample: rows beyond hcut = mean - 1.5 sigma of a
                // random pair are not counted (15x fewer shared-memory atomics); if the tail holds fewer than k sample rows
                // the bootstrap simply yields no bound and the scan starts from its threshold feedback
                const uint16_t* rk = p.rank_tab + m * 257;
#pragma unroll
                for (int rr = 0; rr < 4; rr++)
                    if (d[j][rr] <= hmax && row0 + rr < seg_n) atomicAdd(&s_hist[q * p.R + __ldg(&rk[d[j][rr]])], 1u);
            } else {
                ScanParams sp{};   // the fields emit_group reads
                sp.cand_cnt = p.cand_cnt; sp.hist = p.hist; sp.cand = p.cand; sp.overflow = p.overflow; sp.C = p.C; sp.R = p.R;
                sp.g_world = 0;
                emit_group(sp, q, hmax, d[j][0], d[j][1], d[j][2], d[j][3], seg, row0, seg_n, p.rank_tab + m * 257, dirty_dummy);
                dirty_bits |= 1u << q;
            }
        }
    }
}

// d* = smallest rank whose cumulative count over hq[0..tmax] reaches k, with the counts below / up to it; one warp,
// all bins fetched at once (R <= 1024: 32 per lane). Fewer than k rows: d* = tmax, both counts = all rows.
__device__ __forceinline__ void small_dstar(const uint32_t* hq, uint32_t tmax, uint32_t k, uint32_t lane, uint32_t& dstar, uint32_t& count_lt,
                                            uint32_t& total_le) {
    const uint32_t nb = tmax + 1;
    const uint32_t chunk = (nb + 31) / 32;
    uint32_t v[kSmallMaxR / 32];
    uint32_t sum = 0;
#pragma unroll
    for (uint32_t j = 0; j < kSmallMaxR / 32; j++) {
        const uint32_t r = lane * chunk + j;
        v[j] = (j < chunk && r < nb) ? __ldcg(&hq[r]) : 0u;
        sum += v[j];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    const uint32_t all = __shfl_sync(0xffffffffu, incl, 31);
    const unsigned hit = __ballot_sync(0xffffffffu, incl >= k);
    if (!hit) { dstar = tmax; count_lt = all; total_le = all; return; }
    const int L = __ffs(hit) - 1;
    uint32_t res = 0, lt = 0, le = 0;
    if ((int)lane == L) {
        uint32_t cum = incl - sum;
        bool found = false;
#pragma unroll
        for (uint32_t j = 0; j < kSmallMaxR / 32; j++) {
            if (!found && cum + v[j] >= k) { res = lane * chunk + j; lt = cum; le = cum + v[j]; found = true; }
            cum += v[j];
        }
    }
    dstar = __shfl_sync(0xffffffffu, res, L);
    count_lt = __shfl_sync(0xffffffffu, lt, L);
    total_le = __shfl_sync(0xffffffffu, le, L);
}

// Exact cut of query q by one CTA (all kSmallThreads threads), survivors sorted in shared memory (the ring).
// Returns 0 = done, 1 = candidate list overflowed (the host re-scans the query exactly), 2 = the general k_select has to
// take over (more rows within d* than the shared-memory sort holds).
template <class Shared>
__device__ int small_select(const SmallParams& p, Shared& sh, uint32_t q) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint64_t* s_khi = reinterpret_cast<uint64_t*>(&sh.ring[0][0][0]);
    uint64_t* s_klo = s_khi + kSmallSortCap;
    uint64_t* s_cand = s_klo + kSmallSortCap;
    uint32_t* s_perm = reinterpret_cast<uint32_t*>(s_cand + kSmallSortCap);
    uint16_t* s_nb = reinterpret_cast<uint16_t*>(s_perm + kSmallSortCap);
    const uint32_t oq = p.qsrc[q];
    const uint64_t* list = p.cand + (size_t)q * p.C;
    constexpr int PER = 8;   // list entries per thread and pass, fetched before d* is known (the two latencies overlap)

    if (tid == 0) sh.sel[3] = 0;
    const uint32_t n_all = __ldcg(&p.cand_cnt[q]);
    const bool ovf = __ldcg(&p.overflow[q]) != 0;
    const uint32_t n_list = min(n_all, p.C);
    uint64_t c[PER];
#pragma unroll
    for (int u = 0; u < PER; u++) {
        const uint32_t i = u * kSmallThreads + tid;
        c[u] = i < n_list ? __ldcg(&list[i]) : ~0ull;
    }
    if (warp == 0) {
        uint32_t dstar, count_lt, total_le;
        small_dstar(p.hist + (size_t)q * p.R, min(p.tau_init, p.R - 1), p.k, lane, dstar, count_lt, total_le);
        if (lane == 0) { sh.sel[0] = dstar; sh.sel[1] = count_lt; sh.sel[2] = total_le; }
    }
    __syncthreads();
    ISX_STAMP(7);
    const uint32_t dstar = sh.sel[0], total_le = sh.sel[2];
    const uint32_t n_out = min(p.k, total_le);
    if (tid == 0) {
        p.info[4 * q + 0] = ovf ? 1u : 0u;
        p.info[4 * q + 1] = n_all;
        p.info[4 * q + 2] = dstar;
        p.info[4 * q + 3] = total_le;
    }
    if (ovf) {                                    // truncated list: the host re-scans this query exactly (d*, count known)
        if (tid == 0) p.out_cnt[oq] = 0;
        return 1;
    }
    if (total_le > kSmallSortCap) return 2;       // the general select cuts the ties with its radix pass

    // ---- survivors (rank <= d*) -> shared memory, then their keys (one dependent pair of loads per survivor) ----
    for (uint32_t i0 = 0; i0 < n_list; i0 += kSmallThreads * PER) {
        if (i0) {
#pragma unroll
            for (int u = 0; u < PER; u++) {
                const uint32_t i = i0 + u * kSmallThreads + tid;
                c[u] = i < n_list ? __ldcg(&list[i]) : ~0ull;
            }
        }
#pragma unroll
        for (int u = 0; u < PER; u++) {
            if (c[u] == ~0ull || cand_rank(c[u]) > dstar) continue;
            const uint32_t slot = atomicAdd(&sh.sel[3], 1u);
            if (slot < kSmallSortCap) s_cand[slot] = c[u];
        }
    }
    __syncthreads();
    ISX_STAMP(8);
    const uint32_t n_surv = min(sh.sel[3], kSmallSortCap);
    uint32_t P = 1;
    while (P < n_surv) P <<= 1;
    for (uint32_t e = tid; e < P; e += kSmallThreads) {
        if (e < n_surv) {
            const uint64_t cc = s_cand[e];
            const SegDesc* sd = p.segs + cand_seg(cc);
            const uint64_t* khi = sd->khi;
            const uint64_t* klo = sd->klo;
            const uint32_t len = sd->len_bytes;
            s_khi[e] = khi[cand_row(cc)];
            s_klo[e] = (p.key_words == 2) ? klo[cand_row(cc)] : 0ull;
            s_nb[e] = (uint16_t)(8u * min(p.qlen[q], len));
            s_perm[e] = e;
        } else {
            s_perm[e] = 0xffffffffu;
        }
    }
    __syncthreads();
    ISX_STAMP(9);
    // Only n_out <= k rows are written: every row below d* plus the (k - count_lt) smallest keys among the ties AT d*.
    // With many ties (coarse 64-bit distances) the tie group is cut first by an MSB-first radix select over the keys in
    // shared memory, so the sort below only ever sees the winners.
    uint32_t n_win = n_surv;
    if (n_surv > n_out && n_surv > 256) {
        uint32_t need = p.k - sh.sel[1];   // winners among the ties, >= 1
        uint64_t pre_hi = 0, pre_lo = 0;
        const int n_bytes = p.key_words == 2 ? 16 : 8;
        uint32_t* s_hist256 = reinterpret_cast<uint32_t*>(s_nb + kSmallSortCap);
        for (int b = 0; b < n_bytes; b++) {
            for (uint32_t i = tid; i < 256; i += kSmallThreads) s_hist256[i] = 0;
            __syncthreads();
            for (uint32_t e = tid; e < n_surv; e += kSmallThreads) {
                if (cand_rank(s_cand[e]) != dstar) continue;
                const uint64_t hi = s_khi[e], lo = s_klo[e];
                bool match;
                uint32_t digit;
                if (b < 8) {
                    match = (b == 0) || ((hi >> (64 - 8 * b)) == (pre_hi >> (64 - 8 * b)));
                    digit = (uint32_t)(hi >> (56 - 8 * b)) & 0xffu;
                } else {
                    const int bb = b - 8;
                    match = (hi == pre_hi) && ((bb == 0) || ((lo >> (64 - 8 * bb)) == (pre_lo >> (64 - 8 * bb))));
                    digit = (uint32_t)(lo >> (56 - 8 * bb)) & 0xffu;
                }
                if (match) atomicAdd(&s_hist256[digit], 1u);
            }
            __syncthreads();
            if (warp == 0) {   // digit whose cumulative count reaches `need`
                uint32_t v[8], sum = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) { v[j] = s_hist256[lane * 8 + j]; sum += v[j]; }
                uint32_t incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= (uint32_t)o) incl += t;
                }
                const unsigned hit = __ballot_sync(0xffffffffu, incl >= need);
                const int L = hit ? __ffs(hit) - 1 : 31;
                if ((int)lane == L) {
                    uint32_t cum = incl - sum, dsel = 255, before = cum, in_bin = 0;
                    bool found = false;
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        if (!found && cum + v[j] >= need) { dsel = lane * 8 + j; before = cum; in_bin = v[j]; found = true; }
                        cum += v[j];
                    }
                    sh.sel[4] = dsel;
                    sh.sel[5] = need - before;
                    sh.sel[6] = in_bin;
                    sh.sel[7] = 0;
                }
            }
            __syncthreads();
            const uint64_t dsel = sh.sel[4];
            need = sh.sel[5];
            const uint32_t in_bin = sh.sel[6];
            if (b < 8) pre_hi |= dsel << (56 - 8 * b);
            else pre_lo |= dsel << (56 - 8 * (b - 8));
            if (in_bin <= 32 && b + 1 < n_bytes) {
                // the pivot is the need-th smallest of the <= 32 keys left in the selected bin: one warp ranks them directly
                // (uniformly distributed keys get here after the first byte)
                uint64_t* s_few = reinterpret_cast<uint64_t*>(s_hist256 + 256);   // [32][2]
                for (uint32_t e = tid; e < n_surv; e += kSmallThreads) {
                    if (cand_rank(s_cand[e]) != dstar) continue;
                    const uint64_t hi = s_khi[e], lo = s_klo[e];
                    const int nb = b + 1;   // bytes fixed so far
                    const bool match = nb <= 8 ? (hi >> (64 - 8 * nb)) == (pre_hi >> (64 - 8 * nb))
                                               : (hi == pre_hi && (lo >> (64 - 8 * (nb - 8))) == (pre_lo >> (64 - 8 * (nb - 8))));
                    if (match) { const uint32_t slot = atomicAdd(&sh.sel[7], 1u); if (slot < 32) { s_few[2 * slot] = hi; s_few[2 * slot + 1] = lo; } }
                }
                __syncthreads();
                if (warp == 0) {
                    const uint32_t n_few = min(sh.sel[7], 32u);
                    const uint64_t hi = lane < n_few ? s_few[2 * lane] : ~0ull, lo = lane < n_few ? s_few[2 * lane + 1] : ~0ull;
                    uint32_t rank = 0;
                    for (uint32_t f = 0; f < n_few; f++) {
                        const uint64_t fh = s_few[2 * f], fl = s_few[2 * f + 1];
                        rank += (fh < hi || (fh == hi && fl < lo)) ? 1u : 0u;
                    }
                    if (lane < n_few && rank == need - 1) { s_few[64] = hi; s_few[65] = lo; }
                }
                __syncthreads();
                pre_hi = s_few[64];
                pre_lo = s_few[65];
                break;
            }
        }
        ISX_STAMP(12);
        // winners -> front of the arrays (stable order is irrelevant: they are sorted next)
        if (tid == 0) sh.sel[3] = 0;
        __syncthreads();
        uint32_t* s_tmp = s_perm;   // slot list of the winners
        for (uint32_t e = tid; e < n_surv; e += kSmallThreads) {
            const bool tie = cand_rank(s_cand[e]) == dstar;
            const uint64_t hi = s_khi[e], lo = s_klo[e];
            const bool win = !tie || hi < pre_hi || (hi == pre_hi && (p.key_words == 1 || lo <= pre_lo));
            if (win) s_tmp[atomicAdd(&sh.sel[3], 1u)] = e;
        }
        __syncthreads();
        n_win = sh.sel[3];   // == n_out
        // compact in place through registers (n_win <= k <= 2048 <= 8 per thread)
        uint64_t t_c[8], t_h[8], t_l[8];
        uint16_t t_n[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const uint32_t j = tid + u * kSmallThreads;
            if (j < n_win) { const uint32_t e = s_tmp[j]; t_c[u] = s_cand[e]; t_h[u] = s_khi[e]; t_l[u] = s_klo[e]; t_n[u] = s_nb[e]; }
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const uint32_t j = tid + u * kSmallThreads;
            if (j < n_win) { s_cand[j] = t_c[u]; s_khi[j] = t_h[u]; s_klo[j] = t_l[u]; s_nb[j] = t_n[u]; }
        }
        __syncthreads();
        ISX_STAMP(13);
    }
    if (n_win <= 2 * kSmallThreads) {
        // few rows (the usual case): every element counts the elements that order before it - its final position,
        // no barriers (keys are unique, so the order is strict)
        uint32_t my_pos[2] = {0, 0};
        SortKey my[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const uint32_t e = tid + u * kSmallThreads;
            my[u] = e < n_win ? SortKey{cand_rank(s_cand[e]), s_khi[e], s_klo[e]} : SortKey{0, 0, 0};
        }
        if (tid < n_win) {   // whole warps without an element skip the loop
            const bool two = n_win > (uint32_t)kSmallThreads;
            const bool wide = p.key_words == 2;
            auto less = [](uint32_t ar, uint64_t ah, uint64_t al, uint32_t br, uint64_t bh, uint64_t bl) -> uint32_t {   // branch-free (rank, hi, lo) order
                return (uint32_t)((ar < br) | ((ar == br) & ((ah < bh) | ((ah == bh) & (al < bl)))));
            };
#pragma unroll 4
            for (uint32_t f = 0; f < n_win; f++) {
                const uint32_t fr = cand_rank(s_cand[f]);
                const uint64_t fh = s_khi[f], fl = wide ? s_klo[f] : 0ull;
                my_pos[0] += less(fr, fh, fl, my[0].rank, my[0].hi, my[0].lo);
                if (two) my_pos[1] += less(fr, fh, fl, my[1].rank, my[1].hi, my[1].lo);
            }
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const uint32_t e = tid + u * kSmallThreads;
            if (e < n_win) s_perm[my_pos[u]] = e;
        }
        __syncthreads();
    } else {
        uint32_t P2 = 1;
        while (P2 < n_win) P2 <<= 1;
        for (uint32_t e = tid; e < P2; e += kSmallThreads) s_perm[e] = e < n_win ? e : 0xffffffffu;
        __syncthreads();
        for (uint32_t size = 2; size <= P2; size <<= 1) {
            for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                for (uint32_t i = tid; i < (P2 >> 1); i += kSmallThreads) {
                    const uint32_t lo_i = 2 * i - (i & (stride - 1)), hi_i = lo_i + stride;
                    const bool asc = ((lo_i & size) == 0);
                    const uint32_t pa = s_perm[lo_i], pb = s_perm[hi_i];
                    bool b_lt_a;
                    if (pb == 0xffffffffu) b_lt_a = false;
                    else if (pa == 0xffffffffu) b_lt_a = true;
                    else {
                        const SortKey ka{cand_rank(s_cand[pa]), s_khi[pa], s_klo[pa]}, kb{cand_rank(s_cand[pb]), s_khi[pb], s_klo[pb]};
                        b_lt_a = key_less(kb, ka);
                    }
                    if (b_lt_a == asc) { s_perm[lo_i] = pb; s_perm[hi_i] = pa; }
                }
                __syncthreads();
            }
        }
    }
    ISX_STAMP(10);
    const uint32_t n_write = min(n_out, n_win);
    for (uint32_t j = tid; j < p.k; j += kSmallThreads) {
        const size_t o = (size_t)oq * p.k + j;
        if (j < n_write) {
            const uint32_t e = s_perm[j];
            p.out_khi[o] = s_khi[e];
            p.out_klo[o] = s_klo[e];
            p.out_h[o] = (uint16_t)cand_h(s_cand[e]);
            p.out_n[o] = s_nb[e];
        } else {
            p.out_khi[o] = ~0ull; p.out_klo[o] = ~0ull; p.out_h[o] = 0xffffu; p.out_n[o] = 1;
        }
    }
    if (p.out_codes) {
        for (uint32_t j = tid; j < n_write * 8; j += kSmallThreads) {
            const uint32_t e = s_perm[j >> 3], w = j & 7;
            const uint64_t cc = s_cand[e];
            const SegDesc& sd = p.segs[cand_seg(cc)];
            const uint32_t v = (w < sd.words) ? sd.planes[(size_t)w * sd.cap + cand_row(cc)] : 0u;
            reinterpret_cast<uint32_t*>(p.out_codes)[((size_t)oq * p.k + (j >> 3)) * 8 + w] = v;
        }
    }
    if (tid == 0) p.out_cnt[oq] = n_write;
    ISX_STAMP(11);
    return 0;
}

template <int S, int CH>
__global__ void __launch_bounds__(kSmallThreads, 1) k_scan_small(const __grid_constant__ SmallParams p) {
    using SmallShared = SmallSharedT<S, CH>;
    constexpr int kSmallStages = S;
    extern __shared__ __align__(128) uint8_t smem_small[];
    SmallShared& sh = *reinterpret_cast<SmallShared*>(smem_small);
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem_small + sizeof(SmallShared));   // [T][R], phase 0
    __shared__ unsigned char s_dirty_dummy[kSmallT];
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool producer = warp == kSmallConsumers / 32;
    const bool thr_warp = warp == kSmallConsumers / 32 + 1;
    ISX_STAMP(0);

    // ---- set-up: queries, masks, barriers ----
    for (uint32_t i = tid; i < p.T * 8; i += kSmallThreads) {
        const uint32_t q = i >> 3, w = i & 7, L = p.qlen[q];
        uint32_t v = p.d_queries ? p.d_queries[(size_t)p.qsrc[q] * 8 + w] : p.qwords[q][w];
        if (4 * w >= L) v = 0;
        else if (4 * w + 4 > L) v &= (1u << (8 * (L - 4 * w))) - 1u;
        sh.qw[q][w] = v;
    }
    for (uint32_t i = tid; i < p.n_ranges * p.T * 8; i += kSmallThreads) {
        const uint32_t r = i / (p.T * 8), q = (i / 8) % p.T, w = i & 7;
        const uint32_t m = min(p.qlen[q], p.ranges[r].len_bytes);
        sh.qmask[r][q][w] = (4 * w + 4 <= m) ? 0xffffffffu : (4 * w < m) ? ((1u << (8 * (m - 4 * w))) - 1u) : 0u;
        if (w == 0) {
            sh.m[r][q] = m;
            const float mean = 4.0f * (float)m, sigma = sqrtf(2.0f * (float)m);
            sh.hcut[r][q] = (uint32_t)fmaxf(0.0f, mean - 1.5f * sigma);
        }
    }
    for (uint32_t i = tid; i < p.T * p.R; i += kSmallThreads) s_hist[i] = 0;
    if (tid < kSmallT) { sh.tau[tid] = p.tau_init; s_dirty_dummy[tid] = 0; }
    if (tid == 0) {
        for (int s = 0; s < kSmallStages; s++) { mbar_init(&sh.full[s], 1); mbar_init(&sh.empty[s], kSmallConsumers / 32); }
        sh.dirty = 0;
        sh.stop = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // Copy plan of a unit for this lane. The blocks of a unit are consecutive rows of one segment except where a unit
    // straddles two segments, so plane w of the whole unit is ONE contiguous run: lane w < we copies nblk x 4 KB
    // (2 copies of 32 KB for a 64-bit compare); a straddling unit falls back to one copy per (block, plane).
    // bytes == 0: this lane issues nothing. The producer pipelines three steps over consecutive units so that no global
    // latency sits on the issue path: claim (atomic) -> descriptor loads (unit_load) -> plan (unit_plan) -> TMA issue.
    struct Raw { uint4 d0, d1; uint32_t r, nblk, we, bpu; bool valid; };
    struct Plan { const uint32_t* src; uint32_t bytes, chunk, r, nblk, we; uint32_t seg, row0, seg_n; bool meta, valid; };
    auto unit_load = [&](uint32_t u, uint32_t& r) {
        Raw raw{};
        raw.valid = u != 0xffffffffu;
        if (!raw.valid) return raw;
        while (u >= p.ranges[r].unit0 + p.ranges[r].n_units) r++;
        const SmallRange& rg = p.ranges[r];
        const uint32_t b0 = rg.block0 + (u - rg.unit0) * rg.bpu;
        raw.r = r; raw.we = rg.we; raw.bpu = rg.bpu;
        raw.nblk = min(rg.bpu, rg.block0 + rg.n_blocks - b0);
        if (lane < raw.nblk) {   // lane l reads the descriptor of block l (consumed one iteration later)
            const uint4* src = reinterpret_cast<const uint4*>(p.bdesc + b0 + lane);
            raw.d0 = __ldg(src);
            raw.d1 = __ldg(src + 1);
        }
        return raw;
    };
    auto unit_plan = [&](const Raw& raw) {
        Plan pl{};
        pl.valid = raw.valid;
        if (!raw.valid) return pl;
        pl.r = raw.r; pl.we = raw.we; pl.nblk = raw.nblk;
        const uint64_t base = ((uint64_t)raw.d0.y << 32) | raw.d0.x;
        const uint32_t cap = raw.d0.z, seg = raw.d0.w, row0 = raw.d1.x, seg_n = raw.d1.y;
        const uint32_t seg0 = __shfl_sync(0xffffffffu, seg, 0), row00 = __shfl_sync(0xffffffffu, row0, 0);
        const bool ok = lane >= pl.nblk || (seg == seg0 && row0 == row00 + lane * (uint32_t)kRowsPerStep);
        const bool contiguous = __all_sync(0xffffffffu, ok);
        pl.meta = lane < pl.nblk;
        pl.seg = seg; pl.row0 = row0; pl.seg_n = seg_n;
        if (contiguous) {
            const uint64_t base0 = __shfl_sync(0xffffffffu, base, 0);
            const uint32_t cap0 = __shfl_sync(0xffffffffu, cap, 0);
            if (lane < pl.we) { pl.src = reinterpret_cast<const uint32_t*>(base0) + (size_t)lane * cap0; pl.bytes = pl.nblk * kChunkBytes; pl.chunk = lane * raw.bpu; }
        } else {
            const uint32_t jb = lane / pl.we, w = lane % pl.we;   // lane l -> (block l / we, plane l % we)
            const uint64_t basej = __shfl_sync(0xffffffffu, base, jb < pl.nblk ? jb : 0);
            const uint32_t capj = __shfl_sync(0xffffffffu, cap, jb < pl.nblk ? jb : 0);
            if (jb < pl.nblk) { pl.src = reinterpret_cast<const uint32_t*>(basej) + (size_t)w * capj; pl.bytes = kChunkBytes; pl.chunk = w * raw.bpu + jb; }
        }
        return pl;
    };
    // hm[r][q] for all buckets from the current tau (one warp)
    auto refresh_hm = [&]() {
        for (uint32_t i = lane; i < p.n_ranges * p.T; i += 32) {
            const uint32_t r = i / p.T, q = i % p.T;
            sh.hm[r][q] = __ldg(&p.hmax_tab[(size_t)sh.m[r][q] * p.R + min(sh.tau[q], p.R - 1)]);
        }
    };

    ISX_STAMP(1);
    uint32_t it = 0;   // running stage counter (same sequence on the producer and the consumers)
    for (int phase = 0; phase < 2; phase++) {
        const bool sampling = phase == 0;
        if (producer) {
            // lean issue loop: the plan of the next unit is fetched while the current unit's stage is awaited / issued.
            // phase 0: `sample_units` units per CTA spread over the store; phase 1: units claimed from the global counter.
            const uint64_t policy = l2_evict_first_policy();
            const bool hint = p.l2_evict_first != 0;
            uint32_t r = 0, i = 0;
            bool dry = false;   // the counter ran past the last unit: stop claiming
            // raw claim: the atomic's result stays untouched in lane 0 until the next iteration resolves it
            auto claim_issue = [&]() -> uint32_t {
                uint32_t raw = 0xffffffffu;
                if (sampling) { if (i < p.sample_units) raw = (blockIdx.x + i * gridDim.x) * p.sample_stride; }
                else if (p.static_units) raw = blockIdx.x + i * gridDim.x;
                else if (!dry && lane == 0) raw = atomicAdd(p.unit_counter, 1u);
                i++;
                return raw;
            };
            auto claim_resolve = [&](uint32_t raw) -> uint32_t {
                const uint32_t u = __shfl_sync(0xffffffffu, raw, 0);
                if (u >= p.n_units) { dry = true; return 0xffffffffu; }
                return u;
            };
            // prologue: fill the pipeline (claims run two units ahead of the descriptor loads: an atomic under full
            // memory load takes longer than one 64 KB stage)
            uint32_t raw_claim = claim_issue();
            uint32_t u = claim_resolve(raw_claim);
            Raw raw_loaded = unit_load(u, r);
            raw_claim = claim_issue();
            Plan next = unit_plan(raw_loaded);
            u = claim_resolve(raw_claim);
            if (sampling) r = 0;
            raw_loaded = unit_load(u, r);
            raw_claim = claim_issue();
            uint32_t raw_claim2 = claim_issue();
            for (;; it++) {
                const uint32_t stage = it % kSmallStages, par = (it / kSmallStages) & 1;
                const Plan cur = next;
                const bool have = cur.valid;
                if (have) {
                    next = unit_plan(raw_loaded);            // unit i+1: its descriptors were requested last iteration
                    u = claim_resolve(raw_claim);            // unit i+2: its claim was issued two iterations ago
                    if (sampling) r = 0;
                    raw_loaded = unit_load(u, r);
                    raw_claim = raw_claim2;
                    raw_claim2 = claim_issue();              // unit i+4
                }
                mbar_wait(&sh.empty[stage], par ^ 1);
                if (!have) {   // terminator: an empty stage ends the consumers' loop of this phase
                    if (lane == 0) { sh.stage_nblk[stage] = 0; mbar_arrive(&sh.full[stage]); }
                    it++;
                    break;
                }
                if (cur.meta) { sh.stage_seg[stage][lane] = cur.seg; sh.stage_row0[stage][lane] = cur.row0; sh.stage_segn[stage][lane] = cur.seg_n; }
                if (lane == 0) { sh.stage_range[stage] = cur.r; sh.stage_nblk[stage] = cur.nblk; }
                __syncwarp();
                if (lane == 0) mbar_expect_tx(&sh.full[stage], cur.nblk * cur.we * kChunkBytes);
                __syncwarp();
                if (cur.bytes) {
                    if (hint) tma_bulk_g2s(&sh.ring[stage][cur.chunk][0], cur.src, cur.bytes, &sh.full[stage], policy);
                    else tma_bulk_g2s_plain(&sh.ring[stage][cur.chunk][0], cur.src, cur.bytes, &sh.full[stage]);
                }
            }
            if (!sampling && lane == 0) sh.stop = 1;
        } else if (thr_warp) {
            // thresholds: re-tighten the queries that emitted, pick up what the other CTAs found, publish hm[][]
            if (!sampling) {
                while (!sh.stop) {
                    uint32_t dirty = 0;
                    if (lane == 0) dirty = atomicExch(&sh.dirty, 0u);
                    dirty = __shfl_sync(0xffffffffu, dirty, 0);   // warp-uniform: the tighten below uses full-warp shuffles
                    for (uint32_t q = 0; q < p.T; q++) {
                        if (!(dirty >> q & 1)) continue;
                        const uint32_t tcur = min(__ldcg(&p.tau[q]), p.tau_init);
                        const uint32_t t = small_kth_rank(p.hist + (size_t)q * p.R, tcur, p.k, lane);
                        if (lane == 0 && t < tcur) atomicMin(&p.tau[q], t);
                    }
                    bool changed = false;
                    if (lane < p.T) {
                        const uint32_t t = min(__ldcg(&p.tau[lane]), sh.tau[lane]);   // only ever tightens
                        changed = t != sh.tau[lane];
                        sh.tau[lane] = t;
                    }
                    if (__any_sync(0xffffffffu, changed)) { __syncwarp(); refresh_hm(); }
                    __nanosleep(200);
                }
            }
        } else {
            uint32_t dirty_bits = 0;
            for (;; it++) {
                const uint32_t stage = it % kSmallStages, par = (it / kSmallStages) & 1;
                mbar_wait(&sh.full[stage], par);
                const uint32_t r = sh.stage_range[stage], nblk = sh.stage_nblk[stage];
                if (nblk == 0) {   // terminator stage
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sh.empty[stage]);
                    it++;
                    break;
                }
                switch (p.ranges[r].we) {
                    case 1: small_consume<1, S, CH>(p, sh, stage, r, nblk, sampling, s_hist, s_dirty_dummy, dirty_bits); break;
                    case 2: small_consume<2, S, CH>(p, sh, stage, r, nblk, sampling, s_hist, s_dirty_dummy, dirty_bits); break;
                    case 3: small_consume<3, S, CH>(p, sh, stage, r, nblk, sampling, s_hist, s_dirty_dummy, dirty_bits); break;
                    case 4: small_consume<4, S, CH>(p, sh, stage, r, nblk, sampling, s_hist, s_dirty_dummy, dirty_bits); break;
                    case 5: small_consume<5, S, CH>(p, sh, stage, r, nblk, sampling, s_hist, s_dirty_dummy, dirty_bits); break;
                    case 6: small_consume<6, S, CH>(p, sh, stage, r, nblk, sampling, s_hist, s_dirty_dummy, dirty_bits); break;
                    case 7: small_consume<7, S, CH>(p, sh, stage, r, nblk, sampling, s_hist, s_dirty_dummy, dirty_bits); break;
                    default: small_consume<8, S, CH>(p, sh, stage, r, nblk, sampling, s_hist, s_dirty_dummy, dirty_bits); break;
                }
                __syncwarp();
                if (lane == 0) {
                    if (dirty_bits) { atomicOr(&sh.dirty, dirty_bits); }
                    mbar_arrive(&sh.empty[stage]);
                }
                dirty_bits = 0;
            }
        }
        __syncthreads();
        ISX_STAMP(sampling ? 2 : 4);
        if (sampling) {
            // sample histograms -> global, grid barrier, tau = k-th rank of the sample (same value computed by every CTA)
            for (uint32_t i = tid; i < p.T * p.R; i += kSmallThreads) {
                const uint32_t v = s_hist[i];
                if (v) atomicAdd(&p.shist[i], v);
            }
            __threadfence();
            grid.sync();
            if (warp < p.T) {
                const uint32_t t = small_kth_rank(p.shist + (size_t)warp * p.R, min(p.tau_init, p.R - 1), p.k, lane);
                if (lane == 0) {
                    const uint32_t tq = min(t, p.tau_init);
                    sh.tau[warp] = tq;
                    if (blockIdx.x == 0) atomicMin(&p.tau[warp], tq);
                }
            }
            __syncthreads();
            if (warp == 0) refresh_hm();
            __syncthreads();
            ISX_STAMP(3);
        }
    }

    // ---- phase 2: exact selection, CTA q takes query q; the per-query state is left clean for the next search ----
    __threadfence();
    grid.sync();
    ISX_STAMP(5);
    if (blockIdx.x == gridDim.x - 1 && tid == 0) *p.unit_counter = 0;   // every CTA is past its last claim
    for (uint32_t q = blockIdx.x; q < p.T; q += gridDim.x) {
        const int status = small_select(p, sh, q);
        __syncthreads();
        if (status == 2 && tid == 0) p.info[4 * q + 0] = 2;   // the host runs the general select on this query's state
        if (status == 0) {
            for (uint32_t i = tid; i < p.R; i += kSmallThreads) { p.hist[(size_t)q * p.R + i] = 0; p.shist[(size_t)q * p.R + i] = 0; }
            if (tid == 0) { p.tau[q] = 0xffffffffu; p.cand_cnt[q] = 0; p.overflow[q] = 0; }
        }
        __syncthreads();
    }
    ISX_STAMP(6);
}

}  // namespace isx
