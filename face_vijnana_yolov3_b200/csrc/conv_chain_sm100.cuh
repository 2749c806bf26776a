// A CHAIN of conv layers in one persistent launch (sm_100a).
//
// The residual chains of Darknet-53 at 26x26 and 13x13 (and the 26x26 neck) are runs of layers that all use the same kernel
// instance - 256 x 256 CTA-pair tiles, BK = 64, every output stored by TMA - and whose only coupling is "a tile of layer
// i+1 needs the rows of layer i around it".  conv_igemm_kernel already expresses that coupling with per-128-row-block
// completion counters; what is left of the per-layer cost there is the kernel boundary itself: launch, prologue, the first
// operands' latency and the last tile's drain (~6 us per layer, profiles/r01_timeline_v5.log), plus the wave quantisation
// of every layer taken alone.  Here ONE kernel walks the layer table: every role (A producer, B producer, MMA issuer, the
// two epilogue groups, their store warps) simply continues with the next layer's tiles when it is done with the current
// one, gated by the same counters.  The operand rings, the epilogue rings, TMEM and all barrier phases carry across layers
// (one shared-memory carve-up for the whole chain: A slots are slab-sized, a 1x1 layer's 128-row tile just uses part of
// one), so a CTA's producers fetch the next layer's first tile while its epilogue warps still drain the current layer's
// last one, and the tile -> CTA assignment is rotated from layer to layer so that the uneven tile counts average out over
// the chain instead of costing a partial wave per layer.
//
// Which (layer, tile) items a pair works through, and in which order, is a table look-up (ChainWalk): the static rotation, or
// - FVY_CHAIN_SCHED=1 - per-pair lists from a host list schedule over the known tile lengths and row dependencies.
//
// Same arithmetic as conv_igemm_kernel (same MMA shapes, K order and epilogue), so results are bit-identical.
#pragma once

#include <cstddef>

#include "conv_igemm_sm100.cuh"

namespace fvy {

struct alignas(128) ChainLayer {        // tensor maps read by TMA straight from global memory: 64-byte alignment required
    CUtensorMap tmap_a, tmap_b, tmap_res, tmap_out0, tmap_out1;
    ConvParams p;
    const int* res_flags;      // completion counters of the layer that produced the residual rows (nullptr: before the chain)
    int res_expected, res_blocks;
    int rot;                   // rotation of the tile -> CTA-pair assignment
    int chunks;                // 32-channel chunks of a tile that hold real output channels: 8, or Cout / 32 when Cout < 256 (a 128-wide
                               // 1x1 layer rides the 256-wide pair tile: its upper weight rows are TMA zero fill, chunks 4..7 are not drained)
    int bias_n;                // bias entries to stage (Cout padded to 32)
};
static_assert(sizeof(ChainLayer) % 128 == 0 && offsetof(ChainLayer, tmap_b) == 128, "tensor maps of a chain entry must stay aligned");

constexpr int kChainBN = 256, kChainBK = 64;

// One tap (BK = 64 -> four K = 16 MMAs) of a 256 x 256 pair tile.
__device__ __forceinline__ void chain_mma_tap(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t& accum) {
#pragma unroll
    for (int k = 0; k < kChainBK / 16; ++k) {
        umma_bf16_pair(tmem_d, da + 2 * k, db + 2 * k, idesc, accum);
        accum = 1;
    }
}

// The sequence of (layer, tile) work items of one CTA pair.  Two sources: the static rotation (tile = (pair + rot) % n_pairs,
// + n_pairs, ... layer after layer) or a schedule table written by the host (FVY_CHAIN_SCHED=1: a list schedule over the known
// tile durations and row dependencies; entries (layer << 20) | tile, layer-monotonic per pair - which is what keeps the chain
// deadlock-free: a pair only ever waits for items that precede its own in every other pair's list - terminated by -1).
// Every role of a CTA walks the same sequence with its own copy of this cursor.
struct ChainWalk {
    const ChainLayer* chain;
    const int* sched;
    int n_layers, n_pairs, pair, i, li, tile;
    __device__ __forceinline__ ChainWalk(const ChainLayer* c, int nl, int np, int p, const int* s, int stride)
        : chain(c), sched(s != nullptr ? s + (size_t)p * stride : nullptr), n_layers(nl), n_pairs(np), pair(p), i(0), li(-1), tile(-1) {}
    __device__ __forceinline__ int tiles_of(int l) const {
        const ConvParams& p = chain[l].p;
        return ((p.num_m_tiles + 1) / 2) * p.num_n_tiles;
    }
    __device__ __forceinline__ bool next() {
        if (sched != nullptr) {
            const int e = __ldg(sched + i);
            ++i;
            if (e < 0) return false;
            li = e >> 20; tile = e & 0xFFFFF;
            return true;
        }
        if (li < 0) {
            if (n_layers <= 0) return false;
            li = 0; tile = (pair + chain[0].rot) % n_pairs;
        } else {
            tile += n_pairs;
        }
        while (tile >= tiles_of(li)) {
            if (++li >= n_layers) return false;
            tile = (pair + chain[li].rot) % n_pairs;
        }
        return true;
    }
};

__global__ void __launch_bounds__(kThreads, 1)
conv_chain_kernel(const ChainLayer* __restrict__ chain, const int n_layers, const int nb, const int a_stages, const int b_stages,
                  const int* __restrict__ sched, const int sched_stride) {
    constexpr int BN = kChainBN, BK = kChainBK;
    constexpr int kAcc = 2;
    constexpr uint32_t kTmemCols = kAcc * BN;
    constexpr uint32_t kIdesc = make_idesc_bf16(2 * kBlockM, BN);
    constexpr int kChunks = BN / 32;
    constexpr int kABytes = kBlockM * BK * 2;
    constexpr int kBBytes = (BN / 2) * BK * 2;
    constexpr int kSlabBytes = slab_rows<BK>() * BK * 2;
    constexpr int kRowBytes = BK * 2;
    constexpr int kTileM = 2 * kBlockM;
    const uint32_t cta_rank = cluster_ctarank();
    const int n_pairs = (int)(gridDim.x >> 1);
    const int pair = (int)(blockIdx.x >> 1);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + kSmemBarriers);
    uint64_t* a_empty = a_full + kMaxA;
    uint64_t* b_full = a_empty + kMaxA;
    uint64_t* b_empty = b_full + kMaxB;
    uint64_t* tmem_full = b_empty + kMaxB;
    uint64_t* tmem_empty = tmem_full + kMaxAcc;
    uint64_t* ready_all = tmem_empty + kMaxAcc;             // [2][kMaxRing]  staging buffer holds the residual chunk / is free
    uint64_t* staged_all = ready_all + 2 * kMaxRing;        // [2][kMaxRing]  chunk written by the 128 epilogue threads
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(staged_all + 4 * kMaxRing);
    uint8_t* ring_all = smem + kSmemRing;
    uint8_t* a_ring = ring_all + 2 * nb * kChunkBytes;
    uint8_t* b_ring = a_ring + a_stages * kSlabBytes;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    pdl_launch_dependents();
    if (warp == 0) {
        constexpr int kBars = 2 * kMaxA + 2 * kMaxB + 2 * kMaxAcc + 6 * kMaxRing;
        for (int i = lane; i < kBars; i += 32) {
            uint64_t* bar = a_full + i;
            uint32_t count = 1;
            if (bar >= tmem_empty && bar < tmem_empty + kMaxAcc) count = 16;                     // 2 CTAs x 2 groups x 4 warps
            else if (bar >= staged_all && bar < staged_all + 2 * kMaxRing) count = kEpiThreads / 32;
            mbar_init(bar, count);
        }
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc_pair(tmem_ptr, kTmemCols); tmem_relinquish_pair(); }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    if (warp != kBProducerWarp) pdl_wait();              // the chain's first layer reads what the previous kernel wrote

    if (warp == 0) {
        // ===================== A producer =====================
        if (elect_one()) {
            int as_ = 0; uint32_t aph = 0;
            const bool arrives = cta_rank == 0;
            ChainWalk w(chain, n_layers, n_pairs, pair, sched, sched_stride);
            int cur = -1, nnt = 1, gt = 1, kcn = 1, ntaps = 1, a_choff = 0, wexp = 0, wmargin = 0, wblocks = 0, dep_ready = -1;
            uint32_t tx = 0;
            const int* wflags = nullptr;
            const ChainLayer* L = chain;
            int toff[9];
            while (w.next()) {
                if (w.li != cur) {
                    cur = w.li; L = chain + cur;
                    const ConvParams& p = L->p;
                    nnt = p.num_n_tiles; gt = p.gt; kcn = p.k_chunks; ntaps = p.num_taps; a_choff = p.a_choff;
                    tx = 2u * (uint32_t)(p.a_slab != 0 ? kSlabBytes : kABytes);
                    wflags = p.wait_flags; wexp = p.wait_expected; wmargin = p.wait_margin; wblocks = p.wait_blocks;
#pragma unroll
                    for (int i = 0; i < 9; ++i) toff[i] = p.tap_off[i];
                    tma_prefetch_desc(&L->tmap_a);
                    dep_ready = -1;
                }
                const int tile = w.tile;
                const int m0 = (tile / nnt) * kTileM + (int)cta_rank * kBlockM;
                if (wflags != nullptr)
                    wait_blocks_ready(wflags, wexp, max(0, (m0 - wmargin) >> 7), min(wblocks - 1, (m0 + kBlockM - 1 + wmargin) >> 7), dep_ready);
                for (int tap0 = 0; tap0 < ntaps; tap0 += gt)
                    for (int kc = 0; kc < kcn; ++kc) {
                        // slab: one box serves the gt column taps; 1x1: one 128-row tile per K chunk
                        mbar_wait(&a_empty[as_], aph ^ 1, true);
                        if (arrives) mbar_expect_tx(&a_full[as_], tx);
                        tma_load_2d_pair(a_ring + as_ * kSlabBytes, &L->tmap_a, &a_full[as_], a_choff + kc * BK, m0 + toff[tap0]);
                        if (++as_ == a_stages) { as_ = 0; aph ^= 1; }
                    }
            }
        }
    } else if (warp == kBProducerWarp) {
        // ===================== B producer =====================
        if (elect_one()) {
            int bs = 0; uint32_t bph = 0;
            const bool arrives = cta_rank == 0;
            ChainWalk w(chain, n_layers, n_pairs, pair, sched, sched_stride);
            int cur = -1, nnt = 1, gt = 1, kcn = 1, ntaps = 1;
            const ChainLayer* L = chain;
            while (w.next()) {
                if (w.li != cur) {
                    cur = w.li; L = chain + cur;
                    const ConvParams& p = L->p;
                    nnt = p.num_n_tiles; gt = p.gt; kcn = p.k_chunks; ntaps = p.num_taps;
                }
                const int n0 = (w.tile % nnt) * BN + (int)cta_rank * (BN / 2);
                for (int tap0 = 0; tap0 < ntaps; tap0 += gt)
                    for (int kc = 0; kc < kcn; ++kc)
                        for (int t = 0; t < gt; ++t) {
                            mbar_wait(&b_empty[bs], bph ^ 1, true);
                            if (arrives) mbar_expect_tx(&b_full[bs], 2u * kBBytes);
                            tma_load_2d_pair(b_ring + bs * kBBytes, &L->tmap_b, &b_full[bs], ((tap0 + t) * kcn + kc) * BK, n0);
                            if (++bs == b_stages) { bs = 0; bph ^= 1; }
                        }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA) =====================
        if (cta_rank == 0 && elect_one()) {
            int as_ = 0, bs = 0, acc = 0;
            uint32_t aph = 0, bph = 0, acc_ph = 0;
            const uint64_t desc_hi = make_smem_desc<BK>(0);
            const uint32_t a_ring16 = (smem_u32(a_ring) & 0x3FFFF) >> 4, b_ring16 = (smem_u32(b_ring) & 0x3FFFF) >> 4;
            constexpr uint32_t a_slot16 = (uint32_t)kSlabBytes >> 4, b_slot16 = (uint32_t)kBBytes >> 4, row16 = (uint32_t)kRowBytes >> 4;
            ChainWalk w(chain, n_layers, n_pairs, pair, sched, sched_stride);
            int cur = -1, gt = 1, units = 1;
            while (w.next()) {
                if (w.li != cur) {
                    cur = w.li;
                    const ConvParams& p = chain[cur].p;
                    gt = p.gt; units = p.num_taps * p.k_chunks / gt;
                }
                mbar_wait(&tmem_empty[acc], acc_ph ^ 1, true);
                const uint32_t tmem_d = tmem_base + acc * BN;
                uint32_t accum = 0;
#pragma unroll 1
                for (int u = 0; u < units; ++u) {
                    mbar_wait(&a_full[as_], aph, true);
                    const uint64_t da0 = desc_hi | (uint64_t)(a_ring16 + (uint32_t)as_ * a_slot16);
                    for (int t = 0; t < gt; ++t) {       // gt = 3: the slab read t rows further; gt = 1: the tile
                        mbar_wait(&b_full[bs], bph, true);
                        tc_fence_after();
                        chain_mma_tap(tmem_d, da0 + (uint32_t)t * row16, desc_hi | (uint64_t)(b_ring16 + (uint32_t)bs * b_slot16), kIdesc, accum);
                        umma_commit_pair(&b_empty[bs]);
                        if (++bs == b_stages) { bs = 0; bph ^= 1; }
                    }
                    umma_commit_pair(&a_empty[as_]);
                    if (++as_ == a_stages) { as_ = 0; aph ^= 1; }
                }
                umma_commit_pair(&tmem_full[acc]);
                if (++acc == kAcc) { acc = 0; acc_ph ^= 1; }
            }
        }
    } else if (warp < 10) {
        // ===================== epilogue: group g = warps 2+4g .. 5+4g, column-split (chunks g, g+2, ...) =====================
        const int g = (warp - 2) >> 2;
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int et = (threadIdx.x - 64) & (kEpiThreads - 1);
        const uint32_t swz = (uint32_t)((r >> 1) & 3);
        uint64_t* ready = ready_all + g * kMaxRing;
        uint64_t* staged = staged_all + g * kMaxRing;
        const uint32_t ring_u32 = smem_u32(ring_all + g * nb * kChunkBytes);
        float* sbias = reinterpret_cast<float*>(smem + (g == 0 ? kSmemBias : kSmemRowIdx));   // one bias copy per group (1024 floats)
        const uint32_t sbias_u32 = smem_u32(sbias);
        const int bar_id = 1 + g;
        const int m_rank_off = (int)cta_rank * kBlockM;
        int buf = 0; uint32_t buf_ph = 0;
        int acc = 0; uint32_t acc_ph = 0;
        ChainWalk w(chain, n_layers, n_pairs, pair, sched, sched_stride);
        int cur = -1, nnt = 1, nchunks = kChunks;
        bool has_res = false, leaky = false;
        const ConvParams* pp = &chain[0].p;
        while (w.next()) {
            if (w.li != cur) {
                cur = w.li;
                pp = &chain[cur].p;
                nnt = pp->num_n_tiles; has_res = pp->res != nullptr; leaky = pp->leaky != 0;
                nchunks = chain[cur].chunks;
                const int bias_n = chain[cur].bias_n;
                // this layer's bias: the group's own copy (its previous contents were last read in the group's previous tile)
                named_bar_sync(bar_id, kEpiThreads);
                for (int i = et; i < bias_n; i += kEpiThreads) sbias[i] = __ldg(pp->bias + i);
                named_bar_sync(bar_id, kEpiThreads);
            }
            const ConvParams& p = *pp;
            {
                const int tile = w.tile;
                const int mt = tile / nnt;
                const int m0 = mt * kTileM + m_rank_off;
                const int m = m0 + r;
                const int n0 = (tile - mt * nnt) * BN;
                bool valid = m < p.m_total;
                if (valid) {
                    const int img = (int)__umul64hi((unsigned long long)m, p.magic_plane);
                    const int rem = m - img * p.dom_plane;
                    const int y = (int)__umul64hi((unsigned long long)rem, p.magic_w);
                    const int h = y - p.dom_off;
                    const int w = rem - y * p.dom_w - p.dom_off;
                    valid = (unsigned)h < (unsigned)p.H && (unsigned)w < (unsigned)p.W;
                }
                mbar_wait(&tmem_full[acc], acc_ph);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
#pragma unroll 1
                for (int c = g; c < nchunks; c += 2) {
                    const int c0 = c * 32;
                    const uint32_t sbuf = ring_u32 + (uint32_t)buf * kChunkBytes;
                    const uint32_t myslot = sbuf + (uint32_t)r * 64u;
                    uint32_t accv[32];
                    tmem_ld_32x32(taddr + c0, accv);
                    tmem_ld_wait();
                    float v[32];
                    {
                        const uint32_t b4 = sbias_u32 + (uint32_t)(n0 + c0) * 4u;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b = lds128f(b4 + 16u * j);
                            v[4 * j + 0] = __uint_as_float(accv[4 * j + 0]); v[4 * j + 1] = __uint_as_float(accv[4 * j + 1]);
                            v[4 * j + 2] = __uint_as_float(accv[4 * j + 2]); v[4 * j + 3] = __uint_as_float(accv[4 * j + 3]);
                            add2(v[4 * j + 0], v[4 * j + 1], b.x, b.y);
                            add2(v[4 * j + 2], v[4 * j + 3], b.z, b.w);
                        }
                    }
                    if (leaky) {
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            float m0v, m1v;
                            mul2(m0v, m1v, v[j], v[j + 1], 0.1f);
                            v[j] = fmaxf(v[j], m0v); v[j + 1] = fmaxf(v[j + 1], m1v);
                        }
                    }
                    mbar_wait(&ready[buf], buf_ph);          // the residual chunk has landed in the buffer / its previous contents have left
                    if (has_res) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint4 t = lds128(myslot + ((j ^ swz) << 4));
                            add2(v[8 * j + 0], v[8 * j + 1], __uint_as_float(t.x << 16), __uint_as_float(t.x & 0xFFFF0000u));
                            add2(v[8 * j + 2], v[8 * j + 3], __uint_as_float(t.y << 16), __uint_as_float(t.y & 0xFFFF0000u));
                            add2(v[8 * j + 4], v[8 * j + 5], __uint_as_float(t.z << 16), __uint_as_float(t.z & 0xFFFF0000u));
                            add2(v[8 * j + 6], v[8 * j + 7], __uint_as_float(t.w << 16), __uint_as_float(t.w & 0xFFFF0000u));
                        }
                    }
                    if (!valid) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 pk;
                        pk.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
                        pk.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                        pk.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
                        pk.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
                        sts128(myslot + ((j ^ swz) << 4), pk);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&staged[buf]);
                    if (++buf == nb) { buf = 0; buf_ph ^= 1; }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
                if (++acc == kAcc) { acc = 0; acc_ph ^= 1; }
            }
        }
    } else if (warp >= kStoreWarp0 && warp < kStoreWarp0 + 2) {
        // ===================== store warp of group g =====================
        if (elect_one()) {
            const int g = warp - kStoreWarp0;
            uint8_t* ring = ring_all + g * nb * kChunkBytes;
            uint64_t* ready = ready_all + g * kMaxRing;
            uint64_t* staged = staged_all + g * kMaxRing;
            const int m_rank_off = (int)cta_rank * kBlockM;
            // "prepare" cursor: walks this group's chunks of the whole chain in order, one staging buffer after the other; a
            // buffer is prepared for its next chunk by the residual prefetch (layers with a residual) or a plain arrival.
            // A buffer whose residual rows are not complete yet stays OWED (`pending`) and is retried: the store warp must never
            // block here, or two pairs that wait for each other's rows of an earlier layer would both stop storing the tiles
            // the other one needs.
            ChainWalk pw(chain, n_layers, n_pairs, pair, sched, sched_stride);
            bool p_has = pw.next();
            int p_cur = -1, pchunk = g, pbuf = 0, pdep = -1, pending = 0;
            auto try_prepare = [&]() -> bool {
                if (!p_has) { if (++pbuf == nb) pbuf = 0; return true; }      // past the end: nothing to prepare
                const ChainLayer& L = chain[pw.li];
                const ConvParams& p = L.p;
                if (pw.li != p_cur) { p_cur = pw.li; pdep = -1; }
                if (p.res != nullptr) {
                    const int rm0 = (pw.tile / p.num_n_tiles) * kTileM + m_rank_off;
                    if (L.res_flags != nullptr &&
                        !blocks_ready_now(L.res_flags, L.res_expected, rm0 >> 7, min(L.res_blocks - 1, (rm0 + kBlockM - 1) >> 7), pdep))
                        return false;
                    mbar_expect_tx(&ready[pbuf], kChunkBytes);
                    tma_load_2d(ring + pbuf * kChunkBytes, &L.tmap_res, &ready[pbuf],
                                p.res_choff + (pw.tile % p.num_n_tiles) * BN + pchunk * 32, rm0);
                } else {
                    mbar_arrive(&ready[pbuf]);
                }
                if (++pbuf == nb) pbuf = 0;
                if ((pchunk += 2) >= L.chunks) {
                    pchunk = g;
                    p_has = pw.next();
                }
                return true;
            };
            auto service_pending = [&]() { while (pending > 0 && try_prepare()) --pending; };
            pending = nb;                                              // every buffer starts free
            service_pending();
            int buf = 0, prev = -1; uint32_t sph = 0;
            ChainWalk w(chain, n_layers, n_pairs, pair, sched, sched_stride);
            int cur = -1, nnt = 1, ch0 = 0, ch1 = 0, nchunks = kChunks;
            bool o0 = false, o1 = false;
            int* sig = nullptr;
            const ChainLayer* L = chain;
            while (w.next()) {
                if (w.li != cur) {
                    cur = w.li; L = chain + cur;
                    const ConvParams& p = L->p;
                    nnt = p.num_n_tiles;
                    o0 = p.out[0].tma != 0; o1 = p.out[1].tma != 0;
                    ch0 = p.out[0].choff; ch1 = p.out[1].choff;
                    sig = p.sig_flags;
                    nchunks = L->chunks;
                }
                {
                    const int tile = w.tile;
                    const int m0 = (tile / nnt) * kTileM + m_rank_off;
                    const int n0 = (tile % nnt) * BN;
#pragma unroll 1
                    for (int c = g; c < nchunks; c += 2) {
                        for (uint32_t spin = 0; !mbar_test(&staged[buf], sph); ++spin) {      // keep the owed buffers moving while waiting
                            if (pending > 0) service_pending(); else if (spin > 16) __nanosleep(64);
                            if (spin > (1u << 24)) { printf("fvy: chain store warp timed out (block %d)\n", blockIdx.x); __trap(); }
                        }
                        const uint8_t* sbuf = ring + buf * kChunkBytes;
                        if (o0) tma_store_2d(sbuf, &L->tmap_out0, ch0 + n0 + c * 32, m0);
                        if (o1) tma_store_2d(sbuf, &L->tmap_out1, ch1 + n0 + c * 32, m0);
                        bulk_commit();
                        if (prev >= 0) {
                            bulk_wait_read(1);                     // the previous chunk's store has read its buffer: it is owed its next use
                            ++pending;
                            service_pending();
                        }
                        prev = buf;
                        if (++buf == nb) { buf = 0; sph ^= 1; }
                    }
                    if (sig != nullptr) {                          // publish the tile's row block (see conv_igemm_kernel)
                        bulk_wait_complete(0);
                        fence_proxy_async_all();
                        red_release_gpu_add(sig + (m0 >> 7), 1);
                    }
                }
            }
            bulk_wait_all();
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, kTmemCols);
    }
}

}  // namespace fvy
