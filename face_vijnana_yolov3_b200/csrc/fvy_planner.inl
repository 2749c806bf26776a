// Planner: per-layer tile / pipeline plan, activation buffers, tensor maps, cross-layer counters, layer chains and their schedules.
// Included by fvy_api.cu inside no namespace (opens fvy itself).
namespace fvy {

constexpr int kFuseStemDefault = 0;        // FVY_FUSE_STEM: conv_0 + conv_1 in one kernel
constexpr int kCompactDefault = 2;         // FVY_COMPACT: shared-halo geometry of the narrow deep levels (see build_plan)
constexpr int kChain128Default = 0;        // FVY_CHAIN_128: 128-wide 1x1 layers on the 256-wide pair tile (chain membership at 52^2)
constexpr int kChainSchedDefault = 2;      // FVY_CHAIN_SCHED: 0 static rotation inside the chains, 1 host list schedule, 2 list schedule for chains of few tile waves
constexpr int kSchedMaxWaves = 12;

template <int BN, int BK, bool CTA2>
static int launch_conv_t(fvy_handle* h, Layer& L, int grid) {
    auto kern = conv_igemm_kernel<BN, BK, CTA2>;   // max dynamic smem was raised in query_occ_t at plan time
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = L.smem_bytes; cfg.stream = h->stream;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (h->use_pdl) {   // PDL: prologue overlaps the previous layer's tail
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (CTA2) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = 2; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = at; cfg.numAttrs = na;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, L.tmap_a, L.tmap_b, L.tmap_res, L.tmap_out[0], L.tmap_out[1], L.p));
    ++h->launches;
    return FVY_OK;
}

template <int BN, int BK, bool CTA2>
static int query_occ_t(size_t smem, int* occ) {
    auto kern = conv_igemm_kernel<BN, BK, CTA2>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    if constexpr (CTA2) {
        *occ = 1;       // a cluster of two: one CTA per SM by construction
    } else {
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, kThreads, smem));
    }
    return FVY_OK;
}

#define FVY_DISPATCH(BNv, BKv, CALL)                                                  \
    do {                                                                              \
        if (BKv == 64) {                                                              \
            switch (BNv) {                                                            \
                case 32: return CALL(32, 64, false); case 64: return CALL(64, 64, false);           \
                case 128: return CALL(128, 64, false); case 256: return CALL(256, 64, false);       \
            }                                                                         \
        } else {                                                                      \
            switch (BNv) {                                                            \
                case 32: return CALL(32, 32, false); case 64: return CALL(64, 32, false);           \
                case 128: return CALL(128, 32, false); case 256: return CALL(256, 32, false);       \
            }                                                                         \
        }                                                                             \
        return fail(FVY_E_INVALID, "no kernel instance for tile N=%d K=%d", BNv, BKv); \
    } while (0)

static int launch_conv(fvy_handle* h, Layer& L, int grid) {
    if (L.cta2) return L.BN == 256 ? launch_conv_t<256, 64, true>(h, L, grid) : launch_conv_t<128, 64, true>(h, L, grid);
#define CALL(bn, bk, c2) launch_conv_t<bn, bk, c2>(h, L, grid)
    FVY_DISPATCH(L.BN, L.BK, CALL);
#undef CALL
}
static int query_occ(int BN, int BK, bool cta2, size_t smem, int* occ) {
    if (cta2) return BN == 256 ? query_occ_t<256, 64, true>(smem, occ) : query_occ_t<128, 64, true>(smem, occ);
#define CALL(bn, bk, c2) query_occ_t<bn, bk, c2>(smem, occ)
    FVY_DISPATCH(BN, BK, CALL);
#undef CALL
}

struct TensorBufs { __nv_bfloat16* padded = nullptr; __nv_bfloat16* phase = nullptr; };

static int build_plan(fvy_handle* h) {
    const fvy_config& c = h->cfg;
    std::vector<ConvSpec> specs = h->conv_mode ? single_conv_table(h->conv_cin, h->conv_cout, h->conv_k, h->conv_stride)
                                               : (c.head == FVY_HEAD_YOLO3 ? yolo3_table(c.nb_class) : fd6_table(c.bb_info_c_size));
    const int nmax = c.max_batch;
    // which stored forms does each producer need?
    std::map<int, bool> need_padded, need_phase;
    for (const ConvSpec& s : specs) {
        if (s.src >= 0) { if (s.k == 3 && s.stride == 2) need_phase[s.src] = true; else need_padded[s.src] = true; }
        if (s.res >= 0) need_padded[s.res] = true;
    }
    std::map<int, const ConvSpec*> by_idx;
    for (const ConvSpec& s : specs) by_idx[s.idx] = &s;
    std::map<int, TensorBufs> bufs;
    auto HW = [&](int level, int* H, int* W) { *H = c.net_h >> level; *W = c.net_w >> level; };
    auto env_int = [](const char* name, int dflt) { const char* v = getenv(name); return v && *v ? atoi(v) : dflt; };
    // Stored geometry of a level.  Legacy: every image carries its own one-pixel halo, (H+2) x (W+2).  Compact (the narrow deep levels,
    // 26^2 and 13^2 at 416): neighbouring rows and neighbouring images SHARE their halo - row pitch W+1 (the zero pixel that ends row y
    // is the one that starts row y+1), image pitch (H+1)(W+1) (the zero row on top of image n+1 closes image n; past the last image the
    // tensor map's out-of-bounds zero fill does) - so every tap is still a constant row shift of the flattened buffer while the compute
    // domain, which is this geometry, shrinks from 225 to 196 rows per image at 13^2 (-13 % MMA work) and from 784 to 729 at 26^2 (-7 %).
    // Levels whose outputs leave through the 5-D phase view (rows of >= 32 pixels) need an even row pitch: they keep W+2 and only share
    // the halo ROW between images (FVY_COMPACT=2; 52^2: 2916 -> 2862 rows per image).  Level 1 (written by the stem kernels) stays legacy.
    const int phase_min_w = env_int("FVY_TMA_PHASE_MIN_W", 32);
    const int compact_mode = (c.flags & FVY_CFG_NO_COMPACT) ? 0 : env_int("FVY_COMPACT", kCompactDefault);
    auto is_compact = [&](int level) { return compact_mode >= 1 && level >= 2 && (c.net_w >> level) < phase_min_w; };   // shared columns too
    auto share_rows = [&](int level) { return is_compact(level) || (compact_mode >= 2 && level >= 2); };
    auto geom_w = [&](int level) { return (c.net_w >> level) + (is_compact(level) ? 1 : 2); };
    auto geom_plane = [&](int level) { return (long long)((c.net_h >> level) + (share_rows(level) ? 1 : 2)) * geom_w(level); };
    // concat buffers (yolo3 only): A = [up(conv_84) 256 | skip_61 512] at level 4, B = [up(conv_96) 128 | skip_36 256] at level 3
    __nv_bfloat16 *catA = nullptr, *catB = nullptr;
    if (c.head == FVY_HEAD_YOLO3 && !h->conv_mode) {
        int H, W;
        HW(4, &H, &W);
        if (int e = dev_alloc(h, (void**)&catA, (size_t)nmax * geom_plane(4) * 768 * 2, true)) return e;
        HW(3, &H, &W);
        if (int e = dev_alloc(h, (void**)&catB, (size_t)nmax * geom_plane(3) * 384 * 2, true)) return e;
    }
    // stem: grid of stem_strip_kernel (two pieces per resident warp: 161 vs 166 us) and conv_0's weights in its K order
    {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, stem_strip_kernel<float>, kStripWarps * 32, 0) == cudaSuccess && occ > 0) h->stem_blocks_per_sm = 2 * occ;
        if (const char* sb = getenv("FVY_STEM_BLOCKS")) if (*sb) h->stem_blocks_per_sm = atoi(sb);
    }
    if (int e = dev_alloc(h, (void**)&h->d_stem_w2, 32 * 32 * 2, true)) return e;
    if (h->conv_mode)       // the caller's tensor is packed into this padded buffer (halo = the zeros it is allocated with)
        if (int e = dev_alloc(h, (void**)&h->d_conv_in, (size_t)(h->conv_stride == 2 ? 4 * nmax * geom_plane(1) : nmax * geom_plane(0)) * h->conv_cin * 2, true)) return e;
    // activation buffers
    for (const ConvSpec& s : specs) {
        int H, W;
        HW(s.level, &H, &W);
        TensorBufs tb;
        if (need_padded.count(s.idx))
            if (int e = dev_alloc(h, (void**)&tb.padded, (size_t)nmax * geom_plane(s.level) * s.cout * 2, true)) return e;
        if (need_phase.count(s.idx))      // four planes of the consumer's (next level's) geometry
            if (int e = dev_alloc(h, (void**)&tb.phase, (size_t)4 * nmax * geom_plane(s.level + 1) * s.cout * 2, true)) return e;
        bufs[s.idx] = tb;
    }
    // heads
    int nh = 0;
    for (const ConvSpec& s : specs)
        if (!s.bn) {
            int H, W;
            HW(s.level, &H, &W);
            if (nh >= 3) return fail(FVY_E_INVALID, "more than 3 heads");
            h->gh[nh] = H; h->gw[nh] = W; h->head_c = s.cout;
            if (!h->conv_mode) {      // a single-convolution handle writes to the caller's tensor
                if (int e = dev_alloc(h, (void**)&h->d_logits[nh], (size_t)nmax * H * W * s.cout * 4, true)) return e;
                if (int e = dev_alloc(h, (void**)&h->d_logits_alt[nh], (size_t)nmax * H * W * s.cout * 4, true)) return e;
            }
            ++nh;
        }
    const int bn_cap = c.tile_n_max > 0 ? c.tile_n_max : env_int("FVY_BN", 256);
    const int bn_res_cap = std::min(bn_cap, env_int("FVY_BN_RES", 256));
    const int nb_res = env_int("FVY_NB_RES", 4), nb_plain = env_int("FVY_NB", 3);
    const int stages_cap = env_int("FVY_STAGES", kMaxA);
    const int groups_kn = env_int("FVY_GROUPS_KN", 150);   // (K iterations x 32-column chunks) at or below which a layer gets two epilogue groups
    h->use_pdl = env_int("FVY_PDL", 1) != 0;
    size_t stream_off = 0;
    int head_i = 0;
    for (const ConvSpec& s : specs) {
        Layer L;
        L.s = s;
        HW(s.level, &L.Hout, &L.Wout);
        L.Hin = L.Hout * s.stride; L.Win = L.Wout * s.stride;
        L.stream_off = stream_off;
        stream_off += (size_t)(s.bn ? 4 : 1) * s.cout + (size_t)s.cout * s.cin * s.k * s.k;
        const bool stem = s.src == -1;
        L.taps = stem ? 1 : s.k * s.k;
        L.cin_pad = stem ? 32 : s.cin;
        L.BK = (L.cin_pad % 64 == 0) ? 64 : 32;
        if (L.cin_pad % L.BK) return fail(FVY_E_INVALID, "conv_%d: Cin %d not a multiple of %d", s.idx, s.cin, L.BK);
        L.cout_pad = (s.cout + 31) / 32 * 32;
        if (L.cout_pad > kMaxCout) return fail(FVY_E_INVALID, "conv_%d: Cout %d exceeds %d", s.idx, s.cout, kMaxCout);
        const bool has_res = s.res >= 0;
        int cap_n = has_res ? bn_res_cap : bn_cap;
        if (!stem && s.k == 3 && s.stride == 1 && s.cin >= 128) cap_n = std::min(cap_n, env_int("FVY_BN3", 256));   // deep 3x3 layers: narrower tiles = finer waves
        L.BN = 32;
        for (int bn : {256, 128, 64, 32})
            if (bn <= cap_n && L.cout_pad % bn == 0) { L.BN = bn; break; }
        L.num_n_tiles = L.cout_pad / L.BN;
        // 128-wide 1x1 layers of the residual blocks at 52^2 (conv_13 ... conv_34, conv_101, conv_103): run them on the 256-wide CTA-pair
        // tile (upper 128 weight rows = TMA zero fill, upper output chunks clipped by the TMA store) so that they are eligible for
        // conv_chain_kernel together with the 3x3 layers around them - the kernel boundary costs these 9 us layers more than the
        // doubled (still small) MMA work does.
        const bool wide1x1 = env_int("FVY_CHAIN_128", kChain128Default) != 0 && !(c.flags & FVY_CFG_NO_CHAIN) && !stem && s.k == 1 && s.stride == 1 &&
                             s.bn && s.cout == 128 && s.cin % 64 == 0 && s.src >= 0 && s.idx != 96 && bn_cap >= 256;
        if (wide1x1) { L.BN = 256; L.num_n_tiles = 1; }
        const int gt = L.taps == 9 ? 3 : 1;                       // column taps per filter row
        // CTA pairs (cta_group::2, 256-row tiles, each CTA stages half of the B tile): every 256-wide layer, and the
        // 128-wide 3x3 layers whose whole weight tile then fits in shared memory (conv_5/7/10)
        L.cta2 = !stem && L.BK == 64 && env_int("FVY_CTA2", 1) != 0 &&
                 (L.BN == 256 || (L.BN == 128 && L.taps == 9 && (L.num_n_tiles == 1 ? env_int("FVY_CTA2_128", 1) != 0 : env_int("FVY_CTA2_128", 1) >= 2)) ||
                  (L.BN == 128 && L.taps == 1 && L.num_n_tiles == 1 && env_int("FVY_CTA2_128_1X1", 0) != 0));
        // stride-1 3x3: the three column taps of a filter row read one A slab at row shifts 0, 1, 2
        const bool slab1 = L.taps == 9 && s.stride == 1 && env_int("FVY_SLAB", 1) != 0;
        // stride-2 3x3: column taps 0 and 2 of a filter row are the same input phase one row apart -> one slab for both,
        // a second box for tap 1 (taps are stored in the order s = 0, 2, 1 for these layers)
        const bool slab2 = !stem && L.taps == 9 && s.stride == 2 && env_int("FVY_SLAB2", 1) != 0;
        L.tap_perm = slab2;
        const bool slab = slab1 || slab2;
        const int srows = L.BK == 64 ? slab_rows<64>() : slab_rows<32>();
        const size_t a_tile = (size_t)kBlockM * L.BK * 2, b_tile = (size_t)(L.cta2 ? L.BN / 2 : L.BN) * L.BK * 2;
        const int a_cover = slab ? gt : (L.BK == 32 ? gt : 1);
        const size_t a_slot = slab ? (size_t)(slab2 ? 2 : 1) * srows * L.BK * 2 : a_cover * a_tile;
        int b_cover = (gt == 3 && a_cover == gt && gt * b_tile <= (size_t)env_int("FVY_B3_MAX", 24576)) ? gt : 1;   // a filter row of B tiles per slot (one 3-D TMA box)
        // Layers with a short K loop are epilogue-bound: two epilogue groups alternate tiles.  Deep-K layers keep one
        // group so that the shared memory goes to the operand pipeline instead of a second staging ring.
        const int k_chunks = L.cin_pad / L.BK;
        const int k_iters = (L.taps / (k_chunks == 1 ? gt : 1)) * k_chunks;
        const int groups = env_int("FVY_GROUPS", 0) > 0 ? env_int("FVY_GROUPS", 0) : 2;
        L.deep_k = k_iters * (L.BN / 32) > groups_kn;        // MMA-bound tiles: the epilogue has slack, its latency is what shows
        int nb = has_res ? nb_res : nb_plain, lead = 0;
        nb = std::max(2, std::min(nb, kMaxRing));
        size_t fixed = 1024 + kSmemRing + (size_t)groups * nb * kChunkBytes;
        size_t budget = 232448 - fixed;
        // Resident weights: with a single N tile per CTA the whole [BN, K] weight tile is loaded once and every later
        // tile of the persistent CTA only streams A (half the operand bytes of a 1x1 layer, a quarter of a slab 3x3 layer).
        const size_t b_total = (size_t)L.taps * k_chunks * b_tile;
        int a_stages = 0, b_stages = 0, b_res = 0;
        // conv_5 (stride 2, 64 -> 128): its 74 KB weight half-tile and three 35 KB A slots miss the budget by 6 KB with three
        // staging buffers per group; streaming the weights instead re-reads 72 KB per tile from L2 next to 104 KB of A
        // (~40 B/clk/SM, the MMA issuer starved 67 % of the time) - two staging buffers buy the residency.
        if (slab2 && L.num_n_tiles == 1 && env_int("FVY_RESIDENT", 1) != 0 && b_total + 3 * a_slot > budget && nb > 2 && !has_res &&
            b_total + 3 * a_slot <= budget + (size_t)groups * (nb - 2) * kChunkBytes && env_int("FVY_RESIDENT_NB2", 1) != 0) {
            nb = 2;
            fixed = 1024 + kSmemRing + (size_t)groups * nb * kChunkBytes;
            budget = 232448 - fixed;
        }
        if (L.num_n_tiles == 1 && env_int("FVY_RESIDENT", 1) != 0 && b_total + 3 * a_slot <= budget && !wide1x1) {
            if (L.taps * k_chunks / b_cover > kMaxB && a_cover == gt) b_cover = gt;
            if (L.taps * k_chunks / b_cover <= kMaxB) {
                b_res = 1;
                b_stages = L.taps * k_chunks / b_cover;
                a_stages = (int)std::min<size_t>(std::min(kMaxA, stages_cap), (budget - b_total) / a_slot);
            }
        }
        if (!b_res) {
            // streaming: maximise the taps in flight of the scarcer operand, then the bytes in flight
            const size_t b_slot = b_cover * b_tile;
            // Stride-2 layers: an A slot is two slabs (35 KB at BK = 64) and arrives from the 4-phase planes with the latency of a
            // cold read, while the weights are shared by every CTA and hit L2 - with two A slots the next one can only be requested
            // when the current one is consumed and the MMA issuer starves (FVY_DBG: conv_12 waits for operands 74 % of the time),
            // so these layers take at least three A slots when that leaves two B slots.
            long best = -1;
            for (int min_a = slab2 ? env_int("FVY_S2_MINA", 3) : 2; best < 0 && min_a >= 2; --min_a)
                for (int a = min_a; a <= std::min(kMaxA, stages_cap); ++a)
                    for (int b = 2; b <= std::min(kMaxB, stages_cap * gt); ++b) {
                        const size_t bytes = a * a_slot + b * b_slot;
                        if (bytes > budget) break;
                        const long score = (long)std::min(a * a_cover, b * b_cover) * 1000000 + (long)(bytes >> 10);
                        if (score > best) { best = score; a_stages = a; b_stages = b; }
                    }
            if (best < 0) return fail(FVY_E_INVALID, "conv_%d: no operand pipeline fits in shared memory", s.idx);
        }
        L.stages = a_stages; L.b_stages = b_stages; L.b_resident = b_res;
        L.smem_bytes = fixed + (size_t)a_stages * a_slot + (size_t)b_stages * b_cover * b_tile;
        if (L.smem_bytes > 232448) return fail(FVY_E_INVALID, "conv_%d: shared memory plan %zu exceeds 227 KB", s.idx, L.smem_bytes);
        if (int e = query_occ(L.BN, L.BK, L.cta2, L.smem_bytes, &L.occ)) return e;
        L.occ = 1;   // 320 threads x ~140 registers: one CTA per SM; latency is hidden inside the CTA (stages, two epilogue groups)
        // operands
        const size_t kdim = (size_t)L.taps * L.cin_pad;
        if (int e = dev_alloc(h, (void**)&L.w, (size_t)L.cout_pad * kdim * 2, true)) return e;
        if (int e = dev_alloc(h, (void**)&L.bias, (size_t)std::max(L.cout_pad, L.num_n_tiles * L.BN) * 4, true)) return e;   // zero beyond Cout
        if (b_cover == 3) { if (int e = make_tmap_b3(&L.tmap_b, L.w, L.cin_pad, L.cout_pad, L.taps, L.BK, L.cta2 ? L.BN / 2 : L.BN, 3)) return e; }
        else if (int e = make_tmap_2d(&L.tmap_b, L.w, kdim, L.cout_pad, kdim, L.BK, L.cta2 ? L.BN / 2 : L.BN)) return e;
        ConvParams& p = L.p;
        memset(&p, 0, sizeof(p));
        p.num_taps = L.taps;
        p.k_chunks = L.cin_pad / L.BK;
        p.a_choff = 0;
        p.H = L.Hout; p.W = L.Wout;
        p.leaky = s.leaky ? 1 : 0;
        p.bias = L.bias;
        p.num_n_tiles = L.num_n_tiles;
        p.nb = nb; p.lead = lead; p.epi_groups = groups;
        p.gt = gt; p.a_slab = slab2 ? 2 : (slab1 ? 1 : 0); p.a_cover = a_cover; p.a_stages = a_stages;
        p.b_cover = b_cover; p.b_stages = b_stages; p.b_resident = b_res;
        const void* a_base = nullptr;
        uint64_t a_rows = 0, a_pitch = 0;
        if (stem) {
            a_base = L.w; a_rows = (uint64_t)L.cout_pad; a_pitch = 32;   // placeholder map: conv_0 runs in stem_strip_kernel
            p.dom_plane = L.Hin * L.Win; p.dom_w = L.Win; p.dom_off = 0; p.tap_off[0] = 0;
        } else if (s.stride == 2) {
            // The compute domain is the stored geometry of the OUTPUT level (like a stride-1 layer), and the four phase planes of
            // the input are stored with that same geometry: output pixel (h, w) = domain row m reads phase (r&1, s&1) at position
            // (h + (r>>1), w + (s>>1)) = row m + ((r>>1) - 1) * pitch + ((s>>1) - 1) of that plane: every tap is a constant row
            // shift AND the rows of the domain are the rows of the stored output (TMA stores).
            const long long plane = geom_plane(s.level);
            const int gw = geom_w(s.level);
            a_base = s.src == -4 ? h->d_conv_in : bufs[s.src].phase; a_rows = (uint64_t)(4 * nmax * plane); a_pitch = s.cin;
            p.dom_plane = (int)plane; p.dom_w = gw; p.dom_off = 1;
            for (int r = 0; r < 3; ++r)
                for (int q = 0; q < 3; ++q)
                    p.tap_off[r * 3 + (L.tap_perm ? (q == 0 ? 0 : (q == 2 ? 1 : 2)) : q)] =
                        (int)((((r & 1) << 1) | (q & 1)) * nmax * plane + ((r >> 1) - 1) * gw + ((q >> 1) - 1));
        } else {
            if (s.src == -2) { a_base = catA; a_pitch = 768; }
            else if (s.src == -3) { a_base = catB; a_pitch = 384; }
            else if (s.src == -4) { a_base = h->d_conv_in; a_pitch = s.cin; }
            else { a_base = bufs[s.src].padded; a_pitch = s.cin; }
            const int gw = geom_w(s.level);
            a_rows = (uint64_t)nmax * geom_plane(s.level);
            p.dom_plane = (int)geom_plane(s.level); p.dom_w = gw; p.dom_off = 1;
            if (s.k == 1) p.tap_off[0] = 0;
            else
                for (int r = 0; r < 3; ++r)
                    for (int q = 0; q < 3; ++q) p.tap_off[r * 3 + q] = (r - 1) * gw + (q - 1);
        }
        if (a_base == nullptr) return fail(FVY_E_INVALID, "conv_%d: input buffer missing", s.idx);
        p.magic_plane = ~0ull / (unsigned long long)p.dom_plane + 1ull;
        p.magic_w = ~0ull / (unsigned long long)p.dom_w + 1ull;
        if ((long long)nmax * p.dom_plane >= (1ll << 31)) return fail(FVY_E_INVALID, "conv_%d: %d x %d rows overflow int32", s.idx, nmax, p.dom_plane);
        if (int e = make_tmap_2d(&L.tmap_a, a_base, a_pitch, a_rows, a_pitch, L.BK, slab ? srows : kBlockM)) return e;
        L.tmap_res = L.tmap_a; L.tmap_out[0] = L.tmap_a; L.tmap_out[1] = L.tmap_a;   // placeholders for unused maps
        // rows of the compute domain coincide with rows of the padded (H, W) output buffer (stride-1 convs on a padded input,
        // stride-2 convs on phase planes of the output's geometry)
        const bool coincident = !stem;
        const uint64_t out_rows = (uint64_t)nmax * geom_plane(s.level);
        if (s.res >= 0) {
            p.res = bufs[s.res].padded; p.res_pitch = by_idx[s.res]->cout; p.res_choff = 0;
            if (!p.res) return fail(FVY_E_INVALID, "conv_%d: residual buffer missing", s.idx);
            if (!coincident) return fail(FVY_E_INVALID, "conv_%d: residual on a non stride-1 layer", s.idx);
            if (int e = make_tmap_2d(&L.tmap_res, p.res, p.res_pitch, out_rows, p.res_pitch, 32, kBlockM)) return e;
        }
        // outputs
        int no = 0;
        bool tmap_fail = false;
        const bool use_tma_store = env_int("FVY_TMA_STORE", 1) != 0;
        auto add_out = [&](void* ptr, int kind, int pitch, int choff, int c_real) {
            OutDesc od; od.ptr = ptr; od.aux = nullptr; od.kind = kind; od.pitch = pitch; od.choff = choff; od.nmax = nmax; od.c_real = c_real;
            const int dst_level = kind == OUT_PHASE ? s.level + 1 : (kind == OUT_UP2_PADDED ? s.level - 1 : s.level);
            od.dst_w = geom_w(dst_level); od.dst_plane = (int)geom_plane(dst_level);
            od.tma = (kind == OUT_PADDED && coincident && use_tma_store) ? 1 : 0;
            if (no < 2 && od.tma && make_tmap_2d(&L.tmap_out[no], ptr, (uint64_t)pitch, out_rows, (uint64_t)pitch, 32, kBlockM)) tmap_fail = true;
            // 4-phase form of a stride-1 layer's output: TMA stores through the 5-D phase view (needs an even padded width, which
            // every level with a stride-2 consumer has)
            if (no < 2 && kind == OUT_PHASE && coincident && use_tma_store && env_int("FVY_TMA_PHASE", 1) != 0 && s.stride == 1 && (L.Wout & 1) == 0 &&
                (L.Hout & 1) == 0 && L.Wout >= phase_min_w && !is_compact(s.level)) {   // narrower rows: too many stores per tile (conv_60 @26: 57 vs 56 us)
                CUtensorMap maps[7];
                bool ok = true;
                for (int k = 0; k < 7 && ok; ++k) ok = make_tmap_phase(&maps[k], ptr, nmax, L.Hout, L.Wout, od.dst_w, od.dst_plane, pitch, 1 << k) == FVY_OK;
                void* dmaps = nullptr;
                if (!ok || dev_alloc(h, &dmaps, sizeof(maps), false) || cudaMemcpy(dmaps, maps, sizeof(maps), cudaMemcpyHostToDevice) != cudaSuccess) tmap_fail = true;
                else { od.tma = 2; od.aux = dmaps; L.tmap_out[no] = maps[6]; }
            }
            if (no < 2) p.out[no] = od;
            ++no;
        };
        if (!s.bn) {
            L.head_slot = head_i;
            add_out(h->conv_mode ? (void*)h->d_conv_in /* placeholder: set per call */ : (void*)h->d_logits[head_i], OUT_HEAD_F32, s.cout, 0, s.cout);
            ++head_i;
        } else {
            if (bufs[s.idx].padded) add_out(bufs[s.idx].padded, OUT_PADDED, s.cout, 0, s.cout);
            if (bufs[s.idx].phase) add_out(bufs[s.idx].phase, OUT_PHASE, s.cout, 0, s.cout);
            if (c.head == FVY_HEAD_YOLO3) {
                if (s.idx == 60) add_out(catA, OUT_PADDED, 768, 256, s.cout);
                if (s.idx == 35) add_out(catB, OUT_PADDED, 384, 128, s.cout);
                if (s.idx == 84) add_out(catA, OUT_UP2_PADDED, 768, 0, s.cout);
                if (s.idx == 96) add_out(catB, OUT_UP2_PADDED, 384, 0, s.cout);
            }
        }
        if (tmap_fail) return FVY_E_CUDA;
        if (no == 0) return fail(FVY_E_INVALID, "conv_%d has no consumer", s.idx);
        if (no > 2) return fail(FVY_E_INVALID, "conv_%d has more than two stored forms", s.idx);
        L.primary = p.out[0];
        h->layers.push_back(L);
    }
    h->weight_count = (long long)stream_off;
    // ---- cross-layer tile dependencies: consumer = stride-1 conv reading the plain padded output of a producer whose stored
    // forms all leave by TMA (same geometry: the consumer's compute-domain rows ARE the producer's output rows)
    {
        h->use_flags = env_int("FVY_FLAGS", 1) != 0 && !(c.flags & FVY_CFG_NO_TILE_FLAGS);
        std::map<int, int> layer_of;
        for (size_t i = 0; i < h->layers.size(); ++i) layer_of[h->layers[i].s.idx] = (int)i;
        size_t total = 0;
        for (size_t i = 0; i < h->layers.size() && h->use_flags; ++i) {
            Layer& C = h->layers[i];
            if (C.s.src < 0 || C.s.stride != 1 || !layer_of.count(C.s.src)) continue;
            Layer& P = h->layers[layer_of[C.s.src]];
            if (P.s.level != C.s.level) continue;
            bool pure_tma = true;
            for (int o = 0; o < 2; ++o)
                if (P.p.out[o].kind != OUT_NONE && !P.p.out[o].tma) pure_tma = false;
            if (!pure_tma || P.p.out[0].kind != OUT_PADDED) continue;
            C.wait_on = layer_of[C.s.src];
            P.signals = true;
        }
        for (Layer& L : h->layers)
            if (L.signals) { L.flag_blocks = (int)(((size_t)nmax * L.p.dom_plane + 127) / 128) + 2; total += (size_t)L.flag_blocks; }
        if (total) {
            h->flags_bytes = total * sizeof(int);
            if (int e = dev_alloc(h, (void**)&h->d_flags, h->flags_bytes, true)) return e;
            size_t off = 0;
            for (Layer& L : h->layers)
                if (L.signals) { L.flags = h->d_flags + off; off += (size_t)L.flag_blocks; }
        }
    }
    // ---- chains: runs of consecutive layers of the 256-wide CTA-pair instance, each reading its predecessor's TMA-stored
    // output, go out as ONE persistent launch (conv_chain_kernel)
    {
        h->use_chain = h->use_flags && env_int("FVY_CHAIN", 1) != 0 && !(c.flags & FVY_CFG_NO_CHAIN);
        // 0 = static rotation, 1 = host list schedule, 2 = the list schedule for chains whose layers give a pair at most kSchedMaxWaves tiles
        h->chain_sched = (c.flags & FVY_CFG_CHAIN_SCHED) ? 1 : ((c.flags & FVY_CFG_NO_CHAIN_SCHED) ? 0 : env_int("FVY_CHAIN_SCHED", kChainSchedDefault));
        auto eligible = [&](const Layer& L) {
            if (!L.cta2 || L.BN != 256 || L.BK != 64 || L.s.stride != 1 || L.s.src < 0 || !L.s.bn) return false;
            if (L.p.b_resident || L.p.b_cover != 1) return false;
            if (L.taps == 9 ? L.p.a_slab != 1 : (L.taps != 1 || L.p.a_slab != 0 || L.p.a_cover != 1)) return false;
            for (int o = 0; o < 2; ++o)
                if (L.p.out[o].kind != OUT_NONE && !(L.p.out[o].kind == OUT_PADDED && L.p.out[o].tma)) return false;
            return L.p.out[0].kind == OUT_PADDED;
        };
        const int n = (int)h->layers.size();
        for (int i = 0; i < n && h->use_chain;) {
            if (!eligible(h->layers[i])) { ++i; continue; }
            int j = i + 1;
            const int max_len = env_int("FVY_CHAIN_MAXLEN", 64);
            while (j < n && j - i < max_len && eligible(h->layers[j]) && h->layers[j].wait_on == j - 1 && h->layers[j - 1].signals) ++j;
            if (j - i >= 2) {
                fvy_handle::Chain ch;
                ch.first = i; ch.count = j - i;
                ch.host.resize(ch.count);
                if (int e = dev_alloc(h, (void**)&ch.dev, sizeof(ChainLayer) * ch.count, true)) return e;
                for (int k = i; k < j; ++k) { h->layers[k].chain = (int)h->chains.size(); h->layers[k].chain_pos = k - i; }
                h->chains.push_back(std::move(ch));
            }
            i = j;
        }
        if (!h->chains.empty()) {
            // one shared-memory carve-up for every layer of a chain: slab-sized A slots, per-tap B slots, nb staging buffers per group
            h->chain_nb = env_int("FVY_CHAIN_NB", 3); h->chain_a = env_int("FVY_CHAIN_A", 4); h->chain_b = env_int("FVY_CHAIN_B", 6);
            h->chain_smem = 1024 + kSmemRing + (size_t)2 * h->chain_nb * kChunkBytes + (size_t)h->chain_a * slab_rows<64>() * 64 * 2 +
                            (size_t)h->chain_b * 128 * 64 * 2;
            if (h->chain_smem > 232448) return fail(FVY_E_INVALID, "chain shared-memory plan %zu exceeds 227 KB", h->chain_smem);
            CUDA_TRY(cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
            CUDA_TRY(cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            // The chain kernel's roles spin on tiles produced by other CTAs of the same grid: every pair must be co-resident.  When the
            // device cannot hold num_sms / 2 clusters of this shape at once (MPS / green-context SM limits, MIG), fall back to per-layer launches.
            cudaLaunchConfig_t qc;
            memset(&qc, 0, sizeof(qc));
            qc.gridDim = dim3(h->num_sms & ~1); qc.blockDim = dim3(kThreads); qc.dynamicSmemBytes = h->chain_smem;
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
            qc.attrs = qa; qc.numAttrs = 1;
            int max_clusters = 0;
            if (cudaOccupancyMaxActiveClusters(&max_clusters, conv_chain_kernel, &qc) != cudaSuccess) { cudaGetLastError(); max_clusters = 0; }
            if (max_clusters < (h->num_sms & ~1) / 2) {
                for (Layer& L : h->layers) { L.chain = -1; L.chain_pos = 0; }
                h->chains.clear();
                h->use_chain = false;
            }
        }
    }
    // ---- conv_0 + conv_1 fused (csrc/stem_conv1_fused.cuh): conv_1 = 3x3 / stride 2 / 32 -> 64 reading conv_0, padded output only
    if (h->layers.size() >= 2) {
        const Layer& L0 = h->layers[0];
        Layer& L1 = h->layers[1];
        const bool shape_ok = L0.s.src == -1 && L0.s.cout == 32 && L1.s.src == L0.s.idx && L1.s.k == 3 && L1.s.stride == 2 && L1.s.cin == 32 &&
                              L1.s.cout == 64 && L1.s.bn && L1.s.leaky && L1.s.res < 0 && L1.tap_perm && L1.p.out[0].kind == OUT_PADDED &&
                              L1.p.out[1].kind == OUT_NONE && L1.p.out[0].pitch == 64 && L1.p.out[0].choff == 0;
        h->fuse_stem = shape_ok && env_int("FVY_FUSE_STEM", kFuseStemDefault) != 0 && !(c.flags & FVY_CFG_NO_FUSED_STEM);
        if (h->fuse_stem) {
            if (int e = make_tmap_2d(&h->tmap_w1f, L1.w, 288, 64, 288, 32, 64)) return e;
            const int Ho = c.net_h / 2, Wo = c.net_w / 2;
            FuseParams& f = h->fuse;
            memset(&f, 0, sizeof(f));
            f.H = c.net_h; f.W = c.net_w;
            f.tiles_w = (Wo + kFuseMaxTileW - 1) / kFuseMaxTileW;
            f.tile_w = (Wo + f.tiles_w - 1) / f.tiles_w;
            f.seg_rows = std::max(2, std::min(32, (Ho + 15) / 16));
            if (const int sr = env_int("FVY_FUSE_SEG", 0)) f.seg_rows = std::max(1, sr);
            f.segs = (Ho + f.seg_rows - 1) / f.seg_rows;
            f.w0 = h->d_stem_w2; f.bias0 = L0.bias; f.bias1 = L1.bias; f.out = (__nv_bfloat16*)L1.p.out[0].ptr;
            if ((2 * f.tile_w + 1 + 15) / 16 > kFuseProducers) h->fuse_stem = false;
            CUDA_TRY(cudaFuncSetAttribute(stem_conv1_fused_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuseSmem));
            CUDA_TRY(cudaFuncSetAttribute(stem_conv1_fused_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuseSmem));
            CUDA_TRY(cudaFuncSetAttribute(stem_conv1_fused_kernel<unsigned char>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuseSmem));
        }
    }
    return FVY_OK;
}

// FVY_CHAIN_SCHED=1: a list schedule of the chain's tiles instead of the static rotation.  Every tile is one item of known length
// (taps x K chunks, all of them 256 x 256 x 64 MMAs) that becomes ready when the row blocks of the previous layer its taps reach
// (and the residual rows) are complete; layers are taken in order and, inside a layer, tiles in ascending order, each going to the
// pair that can start it first.  Per pair the list is layer-monotonic, so a pair only ever waits for items that precede its own
// in the other pairs' lists: no cycle.  Times are in units of one tap; `load` and `drain` model the first-operand latency and the
// epilogue + store + counter visibility of a finished tile.
static int schedule_chain(fvy_handle* h, fvy_handle::Chain& ch, int pairs) {
    static const int load = [] { const char* v = getenv("FVY_SCHED_LOAD"); return v && *v ? atoi(v) : 5; }();
    static const int drain = [] { const char* v = getenv("FVY_SCHED_DRAIN"); return v && *v ? atoi(v) : 10; }();
    std::vector<std::vector<int>> lists(pairs);
    std::vector<long long> free_at(pairs, 0);
    std::vector<std::vector<long long>> done(ch.count);
    size_t total = 0;
    for (int k = 0; k < ch.count; ++k) {
        const ConvParams& p = ch.host[k].p;
        const Layer& L = h->layers[ch.first + k];
        const int nnt = p.num_n_tiles, mt_count = (p.num_m_tiles + 1) / 2, tiles = mt_count * nnt;
        const long long dur = (long long)p.num_taps * p.k_chunks;
        if (tiles >= (1 << 20) || k >= (1 << 10)) return fail(FVY_E_INVALID, "chain schedule: %d tiles / layer %d do not fit the entry format", tiles, k);
        done[k].assign(mt_count, 0);
        int res_k = -1;
        if (L.s.res >= 0)
            for (int q = 0; q < k; ++q)
                if (h->layers[ch.first + q].s.idx == L.s.res) res_k = q;
        for (int t = 0; t < tiles; ++t) {
            const int mt = t / nnt;
            long long ready = 0;
            if (k > 0) {
                const long long m0 = (long long)mt * 2 * kBlockM, margin = p.wait_margin;
                const int lo = (int)std::max<long long>(0, (m0 - margin) / (2 * kBlockM));
                const int hi = (int)std::min<long long>((long long)done[k - 1].size() - 1, (m0 + 2 * kBlockM - 1 + margin) / (2 * kBlockM));
                for (int b = lo; b <= hi; ++b) ready = std::max(ready, done[k - 1][b]);
            }
            if (res_k >= 0 && mt < (int)done[res_k].size()) ready = std::max(ready, done[res_k][mt]);
            ready += load;
            int best = 0;
            long long best_start = -1, best_free = -1;
            for (int q = 0; q < pairs; ++q) {
                const long long st = std::max(free_at[q], ready);
                if (best_start < 0 || st < best_start || (st == best_start && free_at[q] > best_free)) { best = q; best_start = st; best_free = free_at[q]; }
            }
            lists[best].push_back((k << 20) | t);
            free_at[best] = best_start + dur;
            done[k][mt] = std::max(done[k][mt], best_start + dur + drain);
            ++total;
        }
    }
    size_t longest = 0;
    for (const auto& l : lists) longest = std::max(longest, l.size());
    const int stride = (int)longest + 1;
    if (ch.d_sched == nullptr || stride > ch.sched_stride) {
        const int alloc_stride = stride + stride / 4 + 8;
        void* pmem = nullptr;
        if (int e = dev_alloc(h, &pmem, (size_t)pairs * alloc_stride * sizeof(int), false)) return e;
        ch.d_sched = (int*)pmem; ch.sched_stride = alloc_stride;
        // the table's address and stride are kernel arguments, i.e. baked into every graph captured so far: drop them
        for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second);
        h->graphs.clear(); h->graph_launches.clear();
    }
    ch.sched_host.assign((size_t)pairs * ch.sched_stride, -1);
    for (int q = 0; q < pairs; ++q) std::copy(lists[q].begin(), lists[q].end(), ch.sched_host.begin() + (size_t)q * ch.sched_stride);
    CUDA_TRY(cudaMemcpyAsync(ch.d_sched, ch.sched_host.data(), ch.sched_host.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    (void)total;
    return FVY_OK;
}

// Per-call chain descriptors (batch-dependent fields), uploaded in stream order before the forward that uses them.
static int prepare_chains(fvy_handle* h, int batch) {
    if (h->chains.empty() || h->chain_batch == batch) return FVY_OK;
    const int pairs = (h->num_sms & ~1) / 2;
    for (fvy_handle::Chain& ch : h->chains) {
        for (int k = 0; k < ch.count; ++k) {
            Layer& L = h->layers[ch.first + k];
            ChainLayer& c = ch.host[k];
            memset(&c, 0, sizeof(c));
            c.tmap_a = L.tmap_a; c.tmap_b = L.tmap_b; c.tmap_res = L.tmap_res; c.tmap_out0 = L.tmap_out[0]; c.tmap_out1 = L.tmap_out[1];
            c.p = L.p;
            c.p.m_total = batch * L.p.dom_plane;
            c.p.num_m_tiles = (c.p.m_total + kBlockM - 1) / kBlockM;
            c.p.epi_groups = 2; c.p.epi_split = 1; c.p.dbg = nullptr;
            c.p.split_from = ((c.p.num_m_tiles + 1) / 2) * c.p.num_n_tiles;
            c.p.sig_flags = L.signals ? L.flags : nullptr;
            c.p.wait_flags = nullptr;
            if (k > 0) {
                const Layer& P = h->layers[L.wait_on];
                c.p.wait_flags = P.flags;
                c.p.wait_expected = P.num_n_tiles * 2;              // both epilogue groups store part of every tile
                c.p.wait_margin = L.s.k == 3 ? L.p.dom_w + 1 : 0;
                c.p.wait_blocks = (batch * P.p.dom_plane + 127) / 128;
            }
            c.res_flags = nullptr;
            if (L.s.res >= 0)
                for (int q = 0; q < k; ++q) {
                    const Layer& R = h->layers[ch.first + q];
                    if (R.s.idx == L.s.res && R.signals) {       // the residual rows come from a layer of this chain
                        c.res_flags = R.flags; c.res_expected = R.num_n_tiles * 2; c.res_blocks = (batch * R.p.dom_plane + 127) / 128;
                    }
                }
            c.rot = (k * 25) % pairs;
            c.chunks = (L.num_n_tiles > 1 || L.cout_pad >= kChainBN) ? kChainBN / 32 : std::max(1, L.cout_pad / 32);
            c.bias_n = L.num_n_tiles * kChainBN;          // the bias array is zero-padded to a whole tile
        }
        CUDA_TRY(cudaMemcpyAsync(ch.dev, ch.host.data(), sizeof(ChainLayer) * ch.count, cudaMemcpyHostToDevice, h->stream));
        // The list schedule pays when a layer is a few tile waves (416 @ batch 40: 3.08 waves at 26^2, 1.68 at 13^2 - the static rotation
        // rounds both up: forward 2.74 -> 2.69 ms); with tens of waves the rotation already balances and the work-list reads only cost
        // (batch 320 @608: 49.5 -> 50.5 ms).
        int most = 0;
        for (int k = 0; k < ch.count; ++k) most = std::max(most, ((ch.host[k].p.num_m_tiles + 1) / 2) * ch.host[k].p.num_n_tiles);
        ch.sched_on = h->chain_sched == 1 || (h->chain_sched >= 2 && most <= kSchedMaxWaves * pairs);
        if (ch.sched_on)
            if (int e = schedule_chain(h, ch, pairs)) return e;
    }
    h->chain_batch = batch;
    return FVY_OK;
}

static int launch_chain(fvy_handle* h, const fvy_handle::Chain& ch) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(h->num_sms & ~1); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = h->chain_smem; cfg.stream = h->stream;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (h->use_pdl) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = 2; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
    ++na;
    cfg.attrs = at; cfg.numAttrs = na;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, conv_chain_kernel, (const ChainLayer*)ch.dev, ch.count, h->chain_nb, h->chain_a, h->chain_b,
                                (const int*)(ch.sched_on ? ch.d_sched : nullptr), ch.sched_stride));
    h->launches += 1;
    return FVY_OK;
}

}  // namespace fvy
