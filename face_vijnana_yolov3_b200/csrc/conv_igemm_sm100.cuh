// Implicit-GEMM convolution for sm_100a: TMA -> shared memory -> tcgen05.mma (accumulators in
// TMEM) -> fused epilogue.  One kernel template serves every conv of the Darknet-53 / YOLOv3
// stack declared by the reference in src/space/yolov3_detect.py:196-311 (_conv_block /
// make_yolov3_model) and the FaceDetector head of src/space/face_detection.py:348-352.
//
// GEMM view:  D[M, N] = A[M, K] * W[N, K]^T,  M = pixels of the compute domain,
// N = Cout, K = taps * Cin.  Activations are NHWC bf16 stored with a one-pixel zero halo
// ("padded" geometry) so that a 3x3 tap is a pure ROW SHIFT of the 2-D matrix [rows, C]:
// the A tile of tap (r,s) is the 128 consecutive rows starting at m0 + (r-1)*(W+2) + (s-1).
// Rows before 0 / past the end are zero-filled by TMA.  Stride-2 convs read a 4-phase
// (space-to-depth by parity) copy written by the producing layer, which again makes every tap a
// row shift.  Halo / out-of-image rows are computed and thrown away (never stored), so the halo
// of every activation buffer keeps the zeros it was allocated with = ZeroPadding2D(1).
//
// Epilogue (all fp32, one rounding per stored tensor): + bias (BatchNorm eps=1e-3 folded on the
// host, yolov3_detect.py:212) -> LeakyReLU(0.1) (:213) -> + residual (:215) -> bf16 store in
// the geometry the consumer wants: padded NHWC, 4-phase (next conv has stride 2), 2x nearest
// up-sampled into a channel slice of a concat buffer (:282-283, :298-299), or dense fp32 head
// logits (:278, :294, :308).
//
// Warp roles (416 threads): warp 0 = A producer, warp 10 = B producer, warp 1 = TMEM allocator + MMA issuer (one elected
// lane each), warps 2..5 / 6..9 = epilogue groups 0 / 1 (TMEM lane quarter = warp_idx & 3), warps 11 / 12 = the store warps
// of the groups (every TMA store and residual prefetch).  Accumulators are 2 (BN = 256) or 4 (BN <= 128) TMEM stages deep.
// Persistent: grid = min(#tiles, #SMs); tiles are strided by gridDim.x.  CTA2 = true: two CTAs of a cluster share one
// 256-row tile (cta_group::2), each staging its own 128 rows of A and half of the B tile.
//
// Operand pipeline: two independent rings.  The K loop runs over (filter row, K chunk, column tap).  The three column taps
// of a filter row read ONE A slab (the activation rows are pixels in row-major padded order, so a column shift is a row
// shift of 1, and tcgen05 shared-memory descriptors swizzle by ADDRESS: any start row inside a TMA-written slab is legal).
// B slots hold one tap or one filter row (a 3-D TMA box), or - resident weights - the whole [BN, K] tile for the CTA's life.
//
// Epilogue data path: the accumulator is drained in chunks of 32 channels.  Each chunk is staged in a ring of 8 KB
// shared-memory buffers laid out exactly as TMA's 64-byte swizzle expects, so that
//   * the residual chunk arrives by TMA (prefetched by the store warp) into the buffer the output chunk will be written to
//     (in place: each thread reads and rewrites its own 64-byte row),
//   * layers whose compute-domain rows ARE rows of the padded output (every layer but the stem) store the chunk with one
//     TMA store per output (halo rows are stored as zeros, which keeps the padding invariant),
//   * the 4-phase form a stride-2 consumer reads leaves by TMA too, through a 5-D view of the phase planes (one box per image row
//     the tile touches, OutDesc::tma == 2); only narrow rows (< 32 pixels) and the 2x up-sampled form are written from the staged
//     chunk with 4 threads per 64-byte row.
// Two groups either split every tile by columns (short drain) or alternate tiles (per-tile set-up amortised).
//
// Layers are chained with programmatic dependent launch; where producer and consumer share a geometry the kernel boundary
// is replaced by per-128-row-block completion counters (ConvParams::sig_flags / wait_flags).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace fvy {

enum OutKind : int {
    OUT_NONE = 0,
    OUT_PADDED = 1,      // [n][H+2][W+2][pitch] bf16, interior only
    OUT_PHASE = 2,       // [4][nmax][H/2+1][W/2+1][pitch] bf16 (consumer is a stride-2 3x3 conv)
    OUT_UP2_PADDED = 3,  // 2x nearest up-sample into [n][2H+2][2W+2][pitch] bf16
    OUT_HEAD_F32 = 4     // [n][H][W][c_real] fp32, dense, no halo
};

struct OutDesc {
    void* ptr;
    const void* aux;   // tma == 2: seven CUtensorMaps (global memory) of the 5-D phase view, box = 1, 2, 4 ... 64 pixel pairs
    int kind;
    int pitch;    // elements per pixel row of the destination buffer
    int choff;    // first destination channel
    int nmax;     // batch capacity of the buffer (phase stride)
    int c_real;   // valid channels (heads: 18 / 255 / 6); others: >= BLOCK_N * num_n_tiles
    int dst_w, dst_plane;   // stored geometry of the destination level: row pitch (pixels) and rows per image (per phase plane)
    int tma;      // 1: stored with TMA (tmap_out[o]); rows of the compute domain coincide with rows of this buffer
                  // 2: OUT_PHASE stored with TMA through the 5-D phase view, one store per image row the tile touches
};

struct ConvParams {
    int num_taps;
    int k_chunks;        // K chunks (of BLOCK_K channels) per tap
    int a_choff;         // first channel of the A buffer
    int tap_off[9];      // row shift per tap (slab: the shift of the slab's first row = tap (g, 0))
    int m_total;         // rows of the compute domain = batch * dom_plane
    int dom_plane;       // rows per image in the compute domain
    int dom_w;           // row pitch (pixels) of the compute domain
    int dom_off;         // 1: padded domain (h = y-1), 0: phase / dense domain (h = y)
    int H, W;            // valid output extent
    int num_m_tiles, num_n_tiles;
    int leaky;
    const float* bias;   // [num_n_tiles * BLOCK_N] fp32
    const __nv_bfloat16* res;   // padded geometry at (H, W), or nullptr
    int res_pitch, res_choff;
    unsigned long long magic_plane, magic_w;   // ceil(2^64 / dom_plane), ceil(2^64 / dom_w): exact division by __umul64hi
    int epi_groups;      // 1 or 2 epilogue warp groups (2: tiles alternate between them)
    int nb;              // staging buffers in EACH group's epilogue ring (2..8)
    int epi_split;       // 1: both groups drain every tile, split by columns (short drain latency); 0: tiles alternate between the groups
    int lead;            // residual prefetch distance in chunks, 2 <= lead <= nb-1
    // Operand pipeline: two independent shared-memory rings.  The K loop runs over (filter row g, K chunk kc, column tap t);
    // an A slot serves a_cover consecutive taps, a B slot b_cover.
    int gt;              // taps per filter row: 3 (3x3) or 1 (1x1, stem)
    int a_slab;          // 1: an A slot is ONE box of slab_rows rows; tap t reads it at a row shift of t (stride-1 3x3 convs)
                         // 2: TWO such boxes (stride-2 3x3 convs, taps ordered s = 0, 2, 1): taps 0 and 1 read the first at row
                         //    shifts 0 and 1 (same input phase), tap 2 reads the second (the other column phase)
    int a_cover;         // taps served by one A slot (slab: gt; separate tiles: 1 or gt)
    int a_stages;
    int b_cover;         // taps per B slot (1 or gt)
    int b_stages;
    int b_resident;      // 1: the B ring holds the whole [BN, K] weight tile of this CTA; loaded once, never released
    int split_from;      // CTA pairs: work items >= split_from are HALF tiles (N/2 columns) of tile split_from + (item - split_from)/2:
                         // the last, partial wave of tiles is spread over twice as many pairs (num_tiles if no split)
    // Cross-layer tile dependencies (instead of the kernel boundary): a layer whose outputs all leave by TMA counts, per
    // 128-row block of its output, the (N tile, epilogue group) pairs whose stores have completed; a consumer of the same
    // geometry starts a tile as soon as the blocks its A rows (and residual rows) come from are complete.
    int* sig_flags;            // this layer's completion counters [ceil(rows / 128)], or nullptr
    const int* wait_flags;     // the producer's counters, or nullptr (then griddepcontrol.wait orders the layers)
    int wait_expected;         // value of a complete counter
    int wait_margin;           // rows before / after the tile that its taps reach (W + 3 for 3x3, 0 for 1x1)
    int wait_blocks;           // number of counters of the producer
    unsigned long long* dbg;   // profiling aid (FVY_DBG): per CTA 8 cycle counters, or nullptr
    OutDesc out[2];
};

// ----------------------------------------------------------------------------------------------
// PTX wrappers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait with back-off.  try_wait suspends the thread in hardware for a short, implementation-defined
// time; the nanosleep keeps a waiting single-thread role from stealing issue slots of the epilogue warp that
// shares its scheduler.  A pipeline bug traps (-> CUDA error at the next sync) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, bool backoff = true) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t spin = 0;; ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred P;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P;\n\t}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if (backoff && spin > 8) __nanosleep(spin > 64 ? 64 : 20);
        if (spin > (1u << 23)) {
            printf("fvy: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, addr, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// smem (swizzled staging chunk) -> global, bulk async-group completion
__device__ __forceinline__ void tma_store_2d(const void* smem_src, const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_5d(const void* smem_src, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most n of this thread's bulk groups still have to READ their shared-memory source
__device__ __forceinline__ void bulk_wait_read(int n) {
    switch (n) {
        case 0: asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); break;
        case 3: asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory"); break;
        case 4: asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); break;
        case 5: asm volatile("cp.async.bulk.wait_group.read 5;" ::: "memory"); break;
        default: asm volatile("cp.async.bulk.wait_group.read 6;" ::: "memory"); break;
    }
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
// programmatic dependent launch
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, issued by ONE thread for the CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
        : "memory");
}
// mbarrier arrives once all MMAs issued so far by this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- CTA-pair (cta_group::2) variants: two CTAs of a cluster cooperate on one 256-row tile --------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory object in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// TMA load issued by either CTA of the pair; the transaction bytes are counted on the LEADER's (rank 0) barrier
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(map_to_cta(smem_u32(bar), 0)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(map_to_cta(smem_u32(bar), 0)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// Remote arrival on the LEADER's barrier.  Relaxed: what the arrival orders is this warp's TMEM reads (complete after
// tcgen05.wait::ld, fenced by tcgen05.fence::before_thread_sync) against the leader's next MMA into the stage - no generic
// memory is handed over.  The .release.cluster form compiles to MEMBAR.ALL.GPU + ERRBAR in front of the arrival, which with
// the epilogue's stores in flight cost ~2000 cycles per tile in every CTA-pair layer (ncu source view / FVY_DBG, conv_7).
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(map_to_cta(smem_u32(bar), 0)) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[128 rows from each CTA's smem] * B[N/2 rows from each CTA's smem]^T, issued by the leader
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum)
        : "memory");
}
// arrives on the barrier at the same offset in BOTH CTAs once the pair's MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}

// 32 lanes x 32 columns of fp32: thread t of the warp gets TMEM lane (base_lane + t), columns [col, col+32).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// Explicit shared-space accesses: the epilogue's pointers are carved out of an aligned uintptr_t, which makes the compiler fall
// back to generic LD/ST (+ two R2UR per access for the memory descriptor); these keep them LDS / STS.
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ int lds32(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// Descriptors (bit layouts: PTX ISA "tcgen05 matrix / instruction descriptor")
// ----------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor for a K-major operand tile whose rows are BK*2 bytes wide and were
// written by TMA with the matching swizzle (BK=64 -> 128B swizzle, BK=32 -> 64B swizzle):
//   [0,14)  start address >> 4        [16,30) leading byte offset >> 4 (ignored for swizzled K-major; 1)
//   [32,46) stride byte offset >> 4 = (8 rows * row bytes) >> 4     [46,48) version = 1 (sm_100)
//   [49,52) base offset = 0 (tile bases are 1024-byte aligned)       [61,64) layout: 2 = SW128, 4 = SW64
template <int BK>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    constexpr uint64_t row_bytes = BK * 2;
    constexpr uint64_t sbo = (8 * row_bytes) >> 4;
    constexpr uint64_t layout = (BK == 64) ? 2 : 4;
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}
// The swizzle XOR is a function of the shared-memory ADDRESS bits (measured on B200, profiles/README.md "slab experiment"):
// a descriptor whose start address lies a few rows into a TMA-written, 1024-byte aligned slab reads exactly the rows
// start + i * row_bytes with the pattern TMA wrote them in, with the base-offset field left 0 (setting it to
// (address >> 7) & 7 double-counts the phase and returns garbage).  That is what lets the three column taps of a filter
// row share one A slab: tap t is the same slab at a start address of t rows further.
template <int BK>
__host__ __device__ constexpr int slab_rows() { return BK == 64 ? 136 : 144; }   // 128 + 2, rounded so that the slab is a multiple of 1024 bytes

// Instruction descriptor, kind::f16: D fp32 (bits 4-5 = 1), A bf16 (bits 7-9 = 1), B bf16 (bits 10-12 = 1),
// A and B K-major (bits 15, 16 = 0), N >> 3 at bits 17-22, M >> 4 at bits 24-28.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ----------------------------------------------------------------------------------------------
// Kernel
// ----------------------------------------------------------------------------------------------
constexpr int kBlockM = 128;
constexpr int kThreads = 416;      // warps: 0 A producer, 1 MMA issuer, 2-5 / 6-9 epilogue groups, 10 B producer, 11 / 12 store warps of the groups
constexpr int kBProducerWarp = 10;
constexpr int kStoreWarp0 = 11;
constexpr int kEpiThreads = 128;
constexpr int kMaxA = 16;       // A ring slots
constexpr int kMaxB = 32;       // B ring slots
constexpr int kMaxRing = 8;
constexpr int kMaxAcc = 4;
constexpr int kChunkBytes = kBlockM * 32 * 2;      // one staged chunk: 128 rows x 32 bf16 = 8 KB
constexpr int kMaxCout = 1024;

// Shared-memory carve-up (offsets from a 1024-byte aligned base)
constexpr int kSmemBarriers = 0;                    // a_full[16] a_empty[16] b_full[32] b_empty[32] tmem_full[4] tmem_empty[4] res_full[2][8] staged[2][8] buf_free[2][8] tmem_ptr (1220 B)
constexpr int kSmemBias = 2048;                     // kMaxCout floats
constexpr int kSmemRowIdx = kSmemBias + kMaxCout * 4;        // int rowidx[2 groups][2 tile parities][2 outputs][128]
constexpr int kSmemRing = kSmemRowIdx + 2 * 2 * 2 * kBlockM * 4; // = 10240, 1024-aligned
static_assert(kSmemRing % 1024 == 0, "staging ring must keep the 512-byte swizzle phase");

template <int BN, int BK>
struct SmemLayout {
    static constexpr int a_bytes = kBlockM * BK * 2;
    static constexpr int b_bytes = BN * BK * 2;
    static constexpr int stage_bytes = a_bytes + b_bytes;
};
__host__ __device__ constexpr int acc_stages(int bn) { return bn <= 128 ? 4 : 2; }   // 512 TMEM columns at BN = 128 and 256

// Packed fp32 pairs (FADD2 / FMUL2 on sm_100): the epilogue is instruction-issue bound, these halve its add / multiply count.
// Each half is an ordinary IEEE round-to-nearest operation - same result as the scalar form.
__device__ __forceinline__ void add2(float& x0, float& x1, float y0, float y1) {
    asm("{\n\t.reg .b64 ra, rb;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tadd.rn.f32x2 ra, ra, rb;\n\tmov.b64 {%0, %1}, ra;\n\t}"
        : "+f"(x0), "+f"(x1) : "f"(y0), "f"(y1));
}
__device__ __forceinline__ void mul2(float& r0, float& r1, float x0, float x1, float c) {
    asm("{\n\t.reg .b64 ra, rb;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\tmul.rn.f32x2 ra, ra, rb;\n\tmov.b64 {%0, %1}, ra;\n\t}"
        : "=f"(r0), "=f"(r1) : "f"(x0), "f"(x1), "f"(c));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// ---- cross-layer tile dependencies ------------------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// Blocks until the producer's 128-row blocks [lo, hi] are complete; `ready` caches the highest block known complete
// (tiles are visited in increasing row order).  The loads that follow are TMA (async proxy): fence after the acquire.
__device__ __forceinline__ void wait_blocks_ready(const int* flags, int expected, int lo, int hi, int& ready) {
    if (hi <= ready) return;
    for (int b = max(lo, ready + 1); b <= hi; ++b) {
        for (uint32_t spin = 0; ld_acquire_gpu(flags + b) < expected; ++spin) {
            __nanosleep(spin < 16 ? 40 : 200);
            if (spin > (1u << 22)) {
                printf("fvy: tile dependency wait timed out (block %d thread %d, counter %d = %d of %d, range %d..%d)\n", blockIdx.x, threadIdx.x, b,
                       ld_acquire_gpu(flags + b), expected, lo, hi);
                __trap();
            }
        }
    }
    ready = hi;
    fence_proxy_async_all();
}
// Non-blocking form: true once blocks [lo, hi] are complete (one acquire load per block not yet known complete).
__device__ __forceinline__ bool blocks_ready_now(const int* flags, int expected, int lo, int hi, int& ready) {
    for (int b = max(lo, ready + 1); b <= hi; ++b) {
        if (ld_acquire_gpu(flags + b) < expected) return false;
        ready = b;
    }
    fence_proxy_async_all();
    return true;
}
// Non-blocking mbarrier phase test.
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void bulk_wait_complete(int n) {   // at most n of this thread's bulk groups still pending (writes performed)
    switch (n) {
        case 0: asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.bulk.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.bulk.wait_group 2;" ::: "memory"); break;
        case 3: asm volatile("cp.async.bulk.wait_group 3;" ::: "memory"); break;
        case 4: asm volatile("cp.async.bulk.wait_group 4;" ::: "memory"); break;
        case 5: asm volatile("cp.async.bulk.wait_group 5;" ::: "memory"); break;
        case 6: asm volatile("cp.async.bulk.wait_group 6;" ::: "memory"); break;
        case 7: asm volatile("cp.async.bulk.wait_group 7;" ::: "memory"); break;
        default: asm volatile("cp.async.bulk.wait_group 8;" ::: "memory"); break;
    }
}

// Work items of the persistent loop: full tiles, then (tail split) the two column halves of each remaining tile.
__device__ __forceinline__ void decode_item(int item, int split_from, int& tile, int& half) {
    if (item < split_from) { tile = item; half = -1; }
    else { const int k = item - split_from; tile = split_from + (k >> 1); half = k & 1; }
}
// First global column (relative to the tile's n0) of 32-column chunk c of a half tile: the pair MMA with N/2 columns takes
// BN/4 weight rows from EACH CTA's B tile, so accumulator columns [0, BN/4) are CTA 0's rows and [BN/4, BN/2) CTA 1's.
template <int BN>
__device__ __forceinline__ int half_chunk_col(int c, int half) {
    constexpr int kQ = BN / 128;                  // chunks per CTA quarter
    return c < kQ ? half * (BN / 4) + c * 32 : BN / 2 + half * (BN / 4) + (c - kQ) * 32;
}


// ---- single-thread producer fast paths: raw shared-memory addresses, no per-iteration address conversion -------------------
__device__ __forceinline__ bool mbar_try_u32(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity, bool backoff) {
    if (mbar_try_u32(bar, parity)) return;
    for (uint32_t spin = 0;; ++spin) {
        if (mbar_try_u32(bar, parity)) return;
        if (backoff && spin > 8) __nanosleep(spin > 64 ? 64 : 20);
        if (spin > (1u << 23)) {
            printf("fvy: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void mbar_expect_tx_u32(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// CTA2: cta_group::2 form, `bar` is the LEADER's barrier (shared::cluster address); else the CTA's own barrier
template <bool CTA2>
__device__ __forceinline__ void tma_load_2d_u32(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    if constexpr (CTA2)
        asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                     "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
                     : "memory");
    else
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                     "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
                     : "memory");
}

// State handed to the MMA issue loop (one thread per CTA / CTA pair).
struct MmaCtx {
    uint64_t *a_full, *a_empty, *b_full, *b_empty, *tmem_full, *tmem_empty;
    uint64_t desc_hi;
    uint32_t a_ring16, b_ring16, a_slot16, b_slot16, b_tap16;   // shared-memory addresses / strides in 16-byte units
    uint32_t a_off16[3];                                        // offset of tap t inside its A slot
    uint32_t tmem_base;
    int a_stages, b_stages, units, first, step, num_tiles, split_from;
    bool bres, bo;
    unsigned long long* dbg;
};

// K loop of one role thread: `units` x GT taps per tile; an A slot serves ACOV taps, a B slot BCOV (compile-time so that the
// per-tap path is a wait, BK/16 MMAs and a commit with no address arithmetic beyond two adds).
template <int BN, int BK, bool CTA2, int GT, int ACOV, int BCOV>
__device__ __forceinline__ void mma_issue(const MmaCtx& c) {
    constexpr int kAcc = acc_stages(BN);
    constexpr uint32_t kIdescFull = make_idesc_bf16(CTA2 ? 2 * kBlockM : kBlockM, BN);
    constexpr uint32_t kIdescHalf = make_idesc_bf16(CTA2 ? 2 * kBlockM : kBlockM, BN / 2);
    constexpr uint32_t kHalfRows16 = (uint32_t)(BN / 4) * (BK * 2) >> 4;     // BN/4 weight rows of this CTA's B tile, in 16-byte units
    int as_ = 0, bs = 0, acc = 0;
    uint32_t aph = 0, bph = 0, acc_ph = 0;
    bool first = true;
    long long dbg_full = 0, dbg_tmem = 0, dbg_taps = 0;
    const long long dbg_t0 = c.dbg ? clock64() : 0;
    for (int item = c.first; item < c.num_tiles; item += c.step) {
        const int half = item < c.split_from ? -1 : ((item - c.split_from) & 1);
        const uint32_t kIdesc = half < 0 ? kIdescFull : kIdescHalf;
        const uint32_t b_half16 = half > 0 ? kHalfRows16 : 0u;
        long long c0 = c.dbg ? clock64() : 0;
        mbar_wait(&c.tmem_empty[acc], acc_ph ^ 1, c.bo);
        if (c.dbg) dbg_tmem += clock64() - c0;
        const uint32_t tmem_d = c.tmem_base + acc * BN;
        uint32_t accum = 0;
        uint64_t da0 = 0, db0 = 0;
#pragma unroll 1
        for (int u = 0; u < c.units; ++u) {
#pragma unroll
            for (int t = 0; t < GT; ++t) {
                if (c.dbg) c0 = clock64();
                if (t % ACOV == 0) {
                    mbar_wait(&c.a_full[as_], aph, c.bo);
                    da0 = c.desc_hi | (uint64_t)(c.a_ring16 + (uint32_t)as_ * c.a_slot16);
                }
                if (t % BCOV == 0) {
                    if (first || !c.bres) mbar_wait(&c.b_full[bs], bph, c.bo);
                    db0 = c.desc_hi | (uint64_t)(c.b_ring16 + (uint32_t)bs * c.b_slot16 + b_half16);
                }
                if (c.dbg) { dbg_full += clock64() - c0; if (first && u == 0 && t == 0) c.dbg[19] = globaltimer_ns(); }
                tc_fence_after();
                const uint64_t da = da0 + c.a_off16[t % ACOV], db = db0 + (uint32_t)(t % BCOV) * c.b_tap16;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    // advance 16 elements (32 bytes) along K inside the swizzle atom: +2 in the >>4 address field
                    if constexpr (CTA2) umma_bf16_pair(tmem_d, da + 2 * k, db + 2 * k, kIdesc, accum);
                    else umma_bf16(tmem_d, da + 2 * k, db + 2 * k, kIdesc, accum);
                    accum = 1;
                }
                if (t % BCOV == BCOV - 1) {
                    if (!c.bres) {                          // frees the slot (in both CTAs of a pair) when these MMAs retire
                        if constexpr (CTA2) umma_commit_pair(&c.b_empty[bs]); else umma_commit(&c.b_empty[bs]);
                    }
                    if (++bs == c.b_stages) { bs = 0; bph ^= 1; }
                }
                if (t % ACOV == ACOV - 1) {
                    if constexpr (CTA2) umma_commit_pair(&c.a_empty[as_]); else umma_commit(&c.a_empty[as_]);
                    if (++as_ == c.a_stages) { as_ = 0; aph ^= 1; }
                }
            }
        }
        // accumulator complete (CTA2: both CTAs' epilogues may drain their half)
        if constexpr (CTA2) umma_commit_pair(&c.tmem_full[acc]); else umma_commit(&c.tmem_full[acc]);
        if (++acc == kAcc) { acc = 0; acc_ph ^= 1; }
        first = false;
        dbg_taps += (long long)c.units * GT;
    }
    if (c.dbg) {
        c.dbg[0] = (unsigned long long)(clock64() - dbg_t0); c.dbg[1] = (unsigned long long)dbg_full; c.dbg[2] = (unsigned long long)dbg_tmem;
        c.dbg[3] = (unsigned long long)dbg_taps;
        c.dbg[20] = globaltimer_ns();
    }
}

template <int BN, int BK, bool CTA2 = false>
__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_res, const __grid_constant__ CUtensorMap tmap_out0,
                  const __grid_constant__ CUtensorMap tmap_out1, const __grid_constant__ ConvParams p) {
    using L = SmemLayout<BN, BK>;
    constexpr int kAcc = acc_stages(BN);
    constexpr uint32_t kTmemCols = kAcc * BN;                      // 128 / 256 / 256 / 512: a power of two >= 32
    // CTA2: the pair computes a 256 x BN tile; each CTA stages its own 128 rows of A and HALF of the B tile (BN/2 rows)
    constexpr uint32_t kIdesc = make_idesc_bf16(CTA2 ? 2 * kBlockM : kBlockM, BN);
    constexpr int kChunks = BN / 32;
    constexpr int kABytes = L::a_bytes;
    constexpr int kBBytes = CTA2 ? L::b_bytes / 2 : L::b_bytes;
    constexpr int kSlabBytes = slab_rows<BK>() * BK * 2;
    constexpr int kRowBytes = BK * 2;
    const uint32_t cta_rank = CTA2 ? cluster_ctarank() : 0u;
    const int cta_step = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;   // tiles advance by the number of clusters
    const int cta_first = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + kSmemBarriers);     // [kMaxA]
    uint64_t* a_empty = a_full + kMaxA;                                        // [kMaxA]
    uint64_t* b_full = a_empty + kMaxA;                                        // [kMaxB]
    uint64_t* b_empty = b_full + kMaxB;                                        // [kMaxB]
    uint64_t* tmem_full = b_empty + kMaxB;                                     // [kMaxAcc]
    uint64_t* tmem_empty = tmem_full + kMaxAcc;                                // [kMaxAcc]
    uint64_t* res_full_all = tmem_empty + kMaxAcc;                             // [2][kMaxRing]
    uint64_t* staged_all = res_full_all + 2 * kMaxRing;                       // [2][kMaxRing]  chunk written by the 128 epilogue threads
    uint64_t* free_all = staged_all + 2 * kMaxRing;                            // [2][kMaxRing]  staging buffer may be overwritten
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(free_all + 2 * kMaxRing);
    float* sbias = reinterpret_cast<float*>(smem + kSmemBias);
    uint8_t* ring_all = smem + kSmemRing;
    const int a_slot_bytes = p.a_slab ? p.a_slab * kSlabBytes : p.a_cover * kABytes;
    const int b_slot_bytes = p.b_cover * kBBytes;
    uint8_t* a_ring = ring_all + p.epi_groups * p.nb * kChunkBytes;
    uint8_t* b_ring = a_ring + p.a_stages * a_slot_bytes;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = (CTA2 ? (p.num_m_tiles + 1) / 2 : p.num_m_tiles) * p.num_n_tiles;   // CTA2: tiles of 256 rows
    const int split_from = CTA2 ? min(p.split_from, num_tiles) : num_tiles;
    const int num_items = num_tiles + (num_tiles - split_from);                                // tail tiles count twice (two halves)

    if (p.m_total < 0) return;    // profiling aid (FVY_NOWORK=2): cost of the bare launch
    if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 32 + 16] = globaltimer_ns();
    pdl_launch_dependents();      // the next layer's CTAs may be scheduled as soon as SMs free up (they wait for our completion below)
    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&tmap_a);
            tma_prefetch_desc(&tmap_b);
            if (p.res != nullptr) tma_prefetch_desc(&tmap_res);
            if (p.out[0].tma == 1) tma_prefetch_desc(&tmap_out0);
            if (p.out[1].tma == 1) tma_prefetch_desc(&tmap_out1);
        }
        // all 152 barriers, 5 per lane (arrival counts: 1, except the accumulator-free and chunk-staged barriers)
        constexpr int kBars = 2 * kMaxA + 2 * kMaxB + 2 * kMaxAcc + 6 * kMaxRing;
        const uint32_t empty_count = (CTA2 ? 8 : 4) * ((BN >= 64 && p.epi_groups == 2 && p.epi_split != 0) ? 2 : 1);
        for (int i = lane; i < kBars; i += 32) {
            uint64_t* bar = a_full + i;
            uint32_t count = 1;
            if (bar >= tmem_empty && bar < tmem_empty + kMaxAcc) count = empty_count;           // every epilogue warp that drains the stage
            else if (bar >= staged_all && bar < staged_all + 2 * kMaxRing) count = kEpiThreads / 32;
            mbar_init(bar, count);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        if constexpr (CTA2) { tmem_alloc_pair(tmem_ptr, kTmemCols); tmem_relinquish_pair(); }
        else { tmem_alloc(tmem_ptr, kTmemCols); tmem_relinquish(); }
    }
    if (warp >= 2)   // bias is a weight, not an activation of the previous layer: safe before griddepcontrol.wait
        for (int i = threadIdx.x - 64; i < p.num_n_tiles * BN; i += kThreads - 64) sbias[i] = __ldg(p.bias + i);
    tc_fence_before();
    __syncthreads();
    if constexpr (CTA2) cluster_sync_all();     // the peer's barriers are initialised before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    // Everything above touched no memory written by the previous layer; from here on we do - except the B producer, which only
    // reads weights: it starts streaming (or loads the resident weight tile) while the previous layer is still finishing.
    if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 32 + 17] = globaltimer_ns();
    if (warp != kBProducerWarp && p.wait_flags == nullptr) pdl_wait();
    if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 32 + 18] = globaltimer_ns();

    // Three single-thread roles feed the tensor pipe: the A producer (warp 0), the B producer (warp 10) and the MMA issuer
    // (warp 1).  Measured on B200 (tools/tma_bench.cu): one thread sustains one TMA instruction per ~170-250 cycles whatever
    // the box size, independent threads scale linearly up to ~72 B/clk/SM with every SM loading - so A and B are issued by
    // different warps, and every loop below keeps running counters instead of divisions.
    const int taps_per_tile = p.num_taps * p.k_chunks;
    const bool bres = p.b_resident != 0;

    if (warp == 0) {
        // ===================== A producer =====================
        if (elect_one()) {
            int as_ = 0; uint32_t aph = 0;
            const bool bo = p.epi_groups == 2;
            const uint32_t tx = (CTA2 ? 2u : 1u) * (uint32_t)a_slot_bytes;   // CTA2: the leader's arrival expects the bytes of BOTH CTAs
            const bool arrives = !CTA2 || cta_rank == 0;
            const int a_loads = p.a_slab ? p.a_slab : p.a_cover;
            const int a_load_bytes = p.a_slab ? kSlabBytes : kABytes;     // smem distance between the boxes of one slot
            const int a_load_tap = p.a_slab == 2 ? 2 : 1;                 // tap index distance between them
            int dep_ready = -1;
            long long dbg_wait = 0;
            // Fast paths.  This loop is ONE thread's scalar code and its length is what a narrow layer's operand rate hangs on
            // (measured on conv_1 / conv_3, ncu source view: ~100 SASS instructions and ~700 cycles per slot in the generic loop
            // below, against ~250 for the TMA instruction itself), so the two common shapes - a slab slot per (filter row, K chunk)
            // with one or two boxes, and one tile box per K chunk of a 1x1 layer - get loops with everything hoisted: row offsets
            // in registers, raw shared-memory addresses advanced by addition, no parameter reloads.
            const bool dbg = p.dbg != nullptr;
            const int fast_mode = split_from < num_tiles ? 0
                                  : (p.a_slab >= 1 && p.gt == 3 && p.a_cover == 3 && p.num_taps == 9) ? p.a_slab
                                  : (p.a_slab == 0 && p.gt == 1 && p.a_cover == 1 && p.num_taps == 1) ? 3 : 0;
            if (fast_mode != 0) {
                const int nnt = p.num_n_tiles, kcn = p.k_chunks, a_choff = p.a_choff, stages = p.a_stages;
                const int* wflags = p.wait_flags;
                const int wexp = p.wait_expected, wmargin = p.wait_margin, wblocks = p.wait_blocks;
                const int n_rows = fast_mode == 3 ? 1 : 3, n_box = fast_mode == 2 ? 2 : 1;
                int off0[3], off1[3];
#pragma unroll
                for (int g = 0; g < 3; ++g) { off0[g] = p.tap_off[fast_mode == 3 ? 0 : 3 * g]; off1[g] = p.tap_off[fast_mode == 3 ? 0 : 3 * g + 2]; }
                const uint32_t ring0 = smem_u32(a_ring), empty0 = smem_u32(a_empty), full_own0 = smem_u32(a_full);
                const uint32_t full_tx0 = CTA2 ? map_to_cta(full_own0, 0) : full_own0;
                const uint32_t empty_end = empty0 + 8u * (uint32_t)stages;
                uint32_t sa = ring0, be = empty0, bf = full_own0, bt = full_tx0, par = 1;     // parity of a free slot's `empty` barrier
                constexpr int kTileRows = CTA2 ? 2 * kBlockM : kBlockM;
                const int rank_off = (int)cta_rank * kBlockM;
                for (int tile = cta_first; tile < num_tiles; tile += cta_step) {
                    const int m0 = (nnt == 1 ? tile : tile / nnt) * kTileRows + rank_off;
                    if (wflags != nullptr)
                        wait_blocks_ready(wflags, wexp, max(0, (m0 - wmargin) >> 7), min(wblocks - 1, (m0 + kBlockM - 1 + wmargin) >> 7), dep_ready);
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        if (g >= n_rows) break;
                        const int row0 = m0 + off0[g], row1 = m0 + off1[g];
                        for (int kc = 0, col = a_choff; kc < kcn; ++kc, col += BK) {
                            const long long c0 = dbg ? clock64() : 0;
                            mbar_wait_u32(be, par, bo);
                            if (dbg) dbg_wait += clock64() - c0;
                            if (arrives) mbar_expect_tx_u32(bf, tx);
                            tma_load_2d_u32<CTA2>(sa, &tmap_a, bt, col, row0);
                            if (n_box == 2) tma_load_2d_u32<CTA2>(sa + (uint32_t)kSlabBytes, &tmap_a, bt, col, row1);
                            sa += (uint32_t)a_slot_bytes; be += 8; bf += 8; bt += 8;
                            if (be == empty_end) { sa = ring0; be = empty0; bf = full_own0; bt = full_tx0; par ^= 1; }
                        }
                    }
                }
            } else
            for (int item = cta_first; item < num_items; item += cta_step) {
                int tile, half;
                decode_item(item, split_from, tile, half);
                const int m0 = (tile / p.num_n_tiles) * (CTA2 ? 2 * kBlockM : kBlockM) + (int)cta_rank * kBlockM;
                if (p.wait_flags != nullptr)      // the producer's row blocks this tile's taps reach
                    wait_blocks_ready(p.wait_flags, p.wait_expected, max(0, (m0 - p.wait_margin) >> 7),
                                      min(p.wait_blocks - 1, (m0 + kBlockM - 1 + p.wait_margin) >> 7), dep_ready);
                for (int tap0 = 0; tap0 < p.num_taps; tap0 += p.gt) {
                    for (int kc = 0; kc < p.k_chunks; ++kc) {
                        for (int t = 0; t < p.gt; t += p.a_cover) {
                            const long long c0 = p.dbg ? clock64() : 0;
                            mbar_wait(&a_empty[as_], aph ^ 1, bo);
                            if (p.dbg) dbg_wait += clock64() - c0;
                            uint8_t* sa = a_ring + as_ * a_slot_bytes;
                            if (arrives) mbar_expect_tx(&a_full[as_], tx);
                            for (int j = 0; j < a_loads; ++j) {
                                if constexpr (CTA2) tma_load_2d_pair(sa + j * a_load_bytes, &tmap_a, &a_full[as_], p.a_choff + kc * BK, m0 + p.tap_off[tap0 + t + j * a_load_tap]);
                                else tma_load_2d(sa + j * a_load_bytes, &tmap_a, &a_full[as_], p.a_choff + kc * BK, m0 + p.tap_off[tap0 + t + j * a_load_tap]);
                            }
                            if (++as_ == p.a_stages) { as_ = 0; aph ^= 1; }
                        }
                    }
                }
            }
            if (p.dbg) p.dbg[blockIdx.x * 32 + 4] = (unsigned long long)dbg_wait;
        }
    } else if (warp == kBProducerWarp) {
        // ===================== B producer (resident weights: the CTA's first tile only) =====================
        if (elect_one()) {
            int bs = 0; uint32_t bph = 0;
            const bool bo = p.epi_groups == 2;
            const uint32_t tx = (CTA2 ? 2u : 1u) * (uint32_t)b_slot_bytes;
            const bool arrives = !CTA2 || cta_rank == 0;
            long long dbg_wait = 0;
            for (int item = cta_first; item < num_items; item += cta_step) {
                int tile, half;
                decode_item(item, split_from, tile, half);      // a half tile loads the whole B tile; the MMA reads its quarter of the rows
                const int n0 = (tile % p.num_n_tiles) * BN + (CTA2 ? (int)cta_rank * (BN / 2) : 0);
                for (int tap0 = 0; tap0 < p.num_taps; tap0 += p.gt) {
                    for (int kc = 0; kc < p.k_chunks; ++kc) {
                        for (int t = 0; t < p.gt; t += p.b_cover) {
                            const long long c0 = p.dbg ? clock64() : 0;
                            if (!bres) mbar_wait(&b_empty[bs], bph ^ 1, bo);
                            if (p.dbg) dbg_wait += clock64() - c0;
                            uint8_t* sb = b_ring + bs * b_slot_bytes;
                            if (arrives) mbar_expect_tx(&b_full[bs], tx);
                            if (p.b_cover == 3) {      // the B tiles of a whole filter row: one 3-D box (Cin chunk, rows, 3 taps)
                                if constexpr (CTA2) tma_load_3d_pair(sb, &tmap_b, &b_full[bs], kc * BK, n0, tap0);
                                else tma_load_3d(sb, &tmap_b, &b_full[bs], kc * BK, n0, tap0);
                            } else {
                                if constexpr (CTA2) tma_load_2d_pair(sb, &tmap_b, &b_full[bs], ((tap0 + t) * p.k_chunks + kc) * BK, n0);
                                else tma_load_2d(sb, &tmap_b, &b_full[bs], ((tap0 + t) * p.k_chunks + kc) * BK, n0);
                            }
                            if (++bs == p.b_stages) { bs = 0; bph ^= 1; }
                        }
                    }
                }
                if (bres) break;
            }
            if (p.dbg) p.dbg[blockIdx.x * 32 + 5] = (unsigned long long)dbg_wait;
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (CTA2: the leader CTA issues for the pair) =====================
        if ((!CTA2 || cta_rank == 0) && elect_one()) {
            MmaCtx c;
            c.a_full = a_full; c.a_empty = a_empty; c.b_full = b_full; c.b_empty = b_empty; c.tmem_full = tmem_full; c.tmem_empty = tmem_empty;
            c.desc_hi = make_smem_desc<BK>(0);
            c.a_ring16 = (smem_u32(a_ring) & 0x3FFFF) >> 4; c.b_ring16 = (smem_u32(b_ring) & 0x3FFFF) >> 4;
            c.a_slot16 = (uint32_t)a_slot_bytes >> 4; c.b_slot16 = (uint32_t)b_slot_bytes >> 4;
            // slab: tap t is the slab read from t rows further (address-based swizzle, see slab_rows)
            if (p.a_slab == 2) { c.a_off16[0] = 0; c.a_off16[1] = (uint32_t)kRowBytes >> 4; c.a_off16[2] = (uint32_t)kSlabBytes >> 4; }
            else for (int t = 0; t < 3; ++t) c.a_off16[t] = (uint32_t)t * ((uint32_t)(p.a_slab ? kRowBytes : kABytes) >> 4);
            c.b_tap16 = (uint32_t)kBBytes >> 4;
            c.a_stages = p.a_stages; c.b_stages = p.b_stages; c.bres = bres; c.bo = p.epi_groups == 2;
            c.tmem_base = tmem_base; c.units = taps_per_tile / p.gt;
            c.first = cta_first; c.step = cta_step; c.num_tiles = num_items; c.split_from = split_from; c.dbg = p.dbg ? p.dbg + blockIdx.x * 32 : nullptr;
            if (p.gt == 1) mma_issue<BN, BK, CTA2, 1, 1, 1>(c);
            else if (p.a_cover == 1) mma_issue<BN, BK, CTA2, 3, 1, 1>(c);
            else if (p.b_cover == 1) mma_issue<BN, BK, CTA2, 3, 3, 1>(c);
            else mma_issue<BN, BK, CTA2, 3, 3, 3>(c);
        }
    } else if (warp < 2 + 4 * p.epi_groups) {
        // ===================== epilogue: group g = warps 2+4g .. 5+4g (128 threads) =====================
        const int g = (warp - 2) >> 2;             // epilogue group
        const int ng = p.epi_groups;
        const int q = warp & 3;                    // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;               // row of the tile handled by this thread (TMEM lane)
        const int et = (threadIdx.x - 64) & (kEpiThreads - 1);   // 0..127 within the group
        const bool has_res = p.res != nullptr;
        // output forms, hoisted out of the tile / chunk loops (uniform per launch)
        const int kind0 = p.out[0].kind, kind1 = p.out[1].kind;
        const bool direct0 = kind0 != OUT_NONE && kind0 != OUT_HEAD_F32 && !p.out[0].tma;
        const bool direct1 = kind1 != OUT_NONE && kind1 != OUT_HEAD_F32 && !p.out[1].tma;
        const bool any_direct = direct0 || direct1;
        const bool head0 = kind0 == OUT_HEAD_F32, head1 = kind1 == OUT_HEAD_F32;
        const int nnt = p.num_n_tiles;
        const int nb = p.nb;
        const uint32_t swz = (uint32_t)((r >> 1) & 3);          // 64-byte swizzle: 16-byte slot j of row r lives at slot j ^ swz
        uint8_t* ring = ring_all + g * nb * kChunkBytes;
        uint64_t* res_full = res_full_all + g * kMaxRing;
        uint64_t* staged = staged_all + g * kMaxRing;
        uint64_t* buf_free = free_all + g * kMaxRing;
        const uint32_t myrow_base = smem_u32(smem + kSmemRowIdx) + (uint32_t)g * 4u * kBlockM * 4u;
        const uint32_t ring_u32 = smem_u32(ring), sbias_u32 = smem_u32(sbias);
        const int bar_id = 1 + g;
        // Two groups: tiles with >= 2 chunks are split by COLUMNS (group g drains chunks g, g+2, ...: half the drain latency per
        // tile, which is what stays exposed at the end of a layer); single-chunk tiles (BN = 32) alternate between the groups.
        const bool split = kChunks >= 2 && ng == 2 && p.epi_split != 0;
        const int tile_step = split ? cta_step : cta_step * ng;
        const int tile_first = split ? cta_first : cta_first + g * cta_step;
        const int c_first = split ? g : 0, c_step = split ? 2 : 1;
        constexpr int kTileM = CTA2 ? 2 * kBlockM : kBlockM;
        const int m_rank_off = (int)cta_rank * kBlockM;

        uint32_t cg = 0;                           // chunks consumed so far by this group
        int buf = 0; uint32_t buf_ph = 0;          // staging buffer of the current chunk and the phase of its barriers
        const bool dbg_on = p.dbg != nullptr && et == 0 && g == 0;
        long long dbg_e_tmem = 0, dbg_e_res = 0, dbg_e_bar = 0, dbg_e_tma = 0, dbg_e_ld = 0, dbg_e_body = 0; const long long dbg_e0 = dbg_on ? clock64() : 0;
        uint32_t it_tile = split ? 0 : g;          // index of the tile in this CTA's sequence (selects the accumulator stage)
        uint32_t my_tiles = 0;
        for (int item = tile_first; item < num_items; item += tile_step, it_tile += (split ? 1 : ng), ++my_tiles) {
            const long long tt0 = dbg_on ? clock64() : 0;
            int tile, half;
            decode_item(item, split_from, tile, half);
            const int n_chunks = half < 0 ? kChunks : kChunks / 2;
            const int as = it_tile % kAcc;
            const uint32_t myrow = myrow_base + (my_tiles & 1) * 2 * kBlockM * 4u;   // double-buffered: a fast thread may be one tile ahead
            const int mt = nnt == 1 ? tile : tile / nnt;
            const int m0 = mt * kTileM + m_rank_off;
            const int m = m0 + r;
            const int n0 = (tile - mt * nnt) * BN;
            // decode the pixel and decide whether the row is a real output (exact division by multiply-high)
            bool valid = m < p.m_total;
            int img = 0, h = 0, w = 0;
            if (valid) {
                img = (int)__umul64hi((unsigned long long)m, p.magic_plane);
                const int rem = m - img * p.dom_plane;
                const int y = (int)__umul64hi((unsigned long long)rem, p.magic_w);
                h = y - p.dom_off;
                w = rem - y * p.dom_w - p.dom_off;
                valid = (unsigned)h < (unsigned)p.H && (unsigned)w < (unsigned)p.W;
            }
            // destination row of this pixel for the outputs that are not stored by TMA
#pragma unroll
            for (int o = 0; o < 2; ++o) {
                if (!(o == 0 ? direct0 : direct1)) continue;
                const OutDesc& od = p.out[o];
                int ridx = -1;
                if (valid) {
                    if (od.kind == OUT_PADDED) ridx = img * od.dst_plane + (h + 1) * od.dst_w + (w + 1);
                    else if (od.kind == OUT_PHASE) {
                        const int hp = h + 1, wp = w + 1;                                // planes of the consumer level's stored geometry
                        ridx = ((((hp & 1) << 1) | (wp & 1)) * od.nmax + img) * od.dst_plane + (hp >> 1) * od.dst_w + (wp >> 1);
                    } else ridx = img * od.dst_plane + (2 * h + 1) * od.dst_w + (2 * w + 1);   // OUT_UP2_PADDED
                }
                sts32(myrow + (uint32_t)(o * kBlockM + r) * 4u, ridx);
            }
            { const long long c0 = dbg_on ? clock64() : 0;
              if (dbg_on) dbg_e_ld += c0 - tt0;        // per-tile set-up (row decode, destination rows)
              mbar_wait(&tmem_full[as], (it_tile / kAcc) & 1);
              if (dbg_on) dbg_e_tmem += clock64() - c0; }
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN;
#pragma unroll 1
            for (int c = c_first; c < n_chunks; c += c_step, ++cg) {
                const int tc = c * 32;                                            // accumulator column of the chunk
                const int c0 = half < 0 ? tc : half_chunk_col<BN>(c, half);       // its first channel inside the N tile
                const uint32_t sbuf = ring_u32 + (uint32_t)buf * kChunkBytes;
                const uint32_t myslot = sbuf + (uint32_t)r * 64u;
                uint32_t acc[32];
                const long long tl0 = dbg_on ? clock64() : 0;
                tmem_ld_32x32(taddr + tc, acc);
                tmem_ld_wait();
                float v[32];
                {
                    const uint32_t b4 = sbias_u32 + (uint32_t)(n0 + c0) * 4u;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 b = lds128f(b4 + 16u * j);
                        v[4 * j + 0] = __uint_as_float(acc[4 * j + 0]); v[4 * j + 1] = __uint_as_float(acc[4 * j + 1]);
                        v[4 * j + 2] = __uint_as_float(acc[4 * j + 2]); v[4 * j + 3] = __uint_as_float(acc[4 * j + 3]);
                        add2(v[4 * j + 0], v[4 * j + 1], b.x, b.y);
                        add2(v[4 * j + 2], v[4 * j + 3], b.z, b.w);
                    }
                }
                if (p.leaky) {   // LeakyReLU(0.1) == max(v, 0.1 v)
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        float m0, m1;
                        mul2(m0, m1, v[j], v[j + 1], 0.1f);
                        v[j] = fmaxf(v[j], m0); v[j + 1] = fmaxf(v[j + 1], m1);
                    }
                }
                {   // the staging buffer is ours once the residual chunk has landed in it / once its previous contents have left
                    const long long tq0 = dbg_on ? clock64() : 0;
                    if (has_res) mbar_wait(&res_full[buf], buf_ph);
                    else mbar_wait(&buf_free[buf], buf_ph ^ 1);
                    if (dbg_on) dbg_e_res += clock64() - tq0;
                }
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
                // fp32 head logits go straight from registers (18 / 255 / 6 valid channels)
#pragma unroll
                for (int o = 0; o < 2; ++o) {
                    const OutDesc& od = p.out[o];
                    if ((o == 0 ? head0 : head1) && valid) {
                        float* dst = reinterpret_cast<float*>(od.ptr) + (((long long)img * p.H + h) * p.W + w) * od.c_real;
                        if ((od.c_real & 3) == 0) {       // wide dense outputs (fvy_conv_run): 16-byte stores
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                if (n0 + c0 + j < od.c_real) *reinterpret_cast<float4*>(dst + n0 + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (n0 + c0 + j < od.c_real) dst[n0 + c0 + j] = v[j];
                        }
                    }
                }
                // stage the bf16 chunk (halo / out-of-image rows as zeros: they ARE the padding of the next layer)
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
                fence_proxy_async();          // generic-proxy writes of the staged chunk -> visible to the async proxy (TMA store)
                if (dbg_on) dbg_e_body += clock64() - tl0;
                if (any_direct) {   // the direct forms read rows staged by other threads
                    const long long tq0 = dbg_on ? clock64() : 0;
                    named_bar_sync(bar_id, kEpiThreads);
                    if (dbg_on) dbg_e_bar += clock64() - tq0;
                }
                // remaining output forms: 4 threads per 64-byte row, 32 rows per pass
                const long long td0 = dbg_on ? clock64() : 0;
#pragma unroll
                for (int o = 0; o < 2; ++o) {
                    if (!(o == 0 ? direct0 : direct1)) continue;
                    const OutDesc& od = p.out[o];
                    __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(od.ptr) + od.choff + n0 + c0 + (et & 3) * 8;
#pragma unroll
                    for (int pass = 0; pass < 4; ++pass) {
                        const int row = pass * 32 + (et >> 2);
                        const int ridx = lds32(myrow + (uint32_t)(o * kBlockM + row) * 4u);
                        if (ridx < 0) continue;
                        const uint4 t = lds128(sbuf + (uint32_t)row * 64u + (uint32_t)(((et & 3) ^ ((row >> 1) & 3)) << 4));
                        if (od.kind == OUT_UP2_PADDED) {
                            const int W2 = od.dst_w;
                            *reinterpret_cast<uint4*>(base + (long long)ridx * od.pitch) = t;
                            *reinterpret_cast<uint4*>(base + (long long)(ridx + 1) * od.pitch) = t;
                            *reinterpret_cast<uint4*>(base + (long long)(ridx + W2) * od.pitch) = t;
                            *reinterpret_cast<uint4*>(base + (long long)(ridx + W2 + 1) * od.pitch) = t;
                        } else {
                            *reinterpret_cast<uint4*>(base + (long long)ridx * od.pitch) = t;
                        }
                    }
                }
                if (dbg_on) dbg_e_tma += clock64() - td0;
                // hand the staged chunk to the store warp
                __syncwarp();
                if (lane == 0) mbar_arrive(&staged[buf]);      // one arrival per warp
                if (++buf == nb) { buf = 0; buf_ph ^= 1; }
            }
            // all TMEM reads of this accumulator stage are complete (tmem_ld_wait above): hand it back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CTA2) mbar_arrive_leader(&tmem_empty[as]);   // the leader's MMA warp waits for both CTAs' epilogues
                else mbar_arrive(&tmem_empty[as]);
            }
        }
        if (dbg_on) {
            unsigned long long* d = p.dbg + blockIdx.x * 32 + 8;
            d[0] = (unsigned long long)(clock64() - dbg_e0); d[1] = (unsigned long long)dbg_e_tmem; d[2] = (unsigned long long)dbg_e_res;
            d[3] = (unsigned long long)dbg_e_bar; d[4] = (unsigned long long)dbg_e_tma; d[5] = cg; d[6] = (unsigned long long)dbg_e_ld; d[7] = (unsigned long long)dbg_e_body;
            p.dbg[blockIdx.x * 32 + 21] = globaltimer_ns();
        }
    } else if (warp >= kStoreWarp0 && warp < kStoreWarp0 + p.epi_groups) {
        // ===================== store warp of epilogue group g: every TMA instruction of the epilogue =====================
        // A TMA instruction costs its issuing thread ~200 cycles; here they overlap the 128 math threads instead of stalling
        // them.  Per staged chunk: TMA store(s) of the chunk, then - once the PREVIOUS chunk's store has read its buffer -
        // that buffer is refilled with the residual of the chunk that will use it next (nb chunks later) or marked free.
        if (elect_one()) {
            const int g = warp - kStoreWarp0;
            const int ng = p.epi_groups, nb = p.nb;
            const bool has_res = p.res != nullptr;
            const bool any_tma = p.out[0].tma || p.out[1].tma;
            uint8_t* ring = ring_all + g * nb * kChunkBytes;
            uint64_t* res_full = res_full_all + g * kMaxRing;
            uint64_t* staged = staged_all + g * kMaxRing;
            uint64_t* buf_free = free_all + g * kMaxRing;
            const bool split = kChunks >= 2 && ng == 2 && p.epi_split != 0;
            const int tile_step = split ? cta_step : cta_step * ng;
            const int tile_first = split ? cta_first : cta_first + g * cta_step;
            const int c_first = split ? g : 0, c_step = split ? 2 : 1;
            constexpr int kTileM = CTA2 ? 2 * kBlockM : kBlockM;
            const int m_rank_off = (int)cta_rank * kBlockM;
            // residual prefetch cursor: walks this group's chunks in order, one staging buffer after the other
            int pf_item = tile_first, pf_chunk = c_first, pf_buf = 0;
            int dep_ready = -1;
            auto prefetch_res = [&]() {
                if (pf_item < num_items) {
                    int tile, half;
                    decode_item(pf_item, split_from, tile, half);
                    if (p.wait_flags != nullptr) {     // the residual rows were written by an earlier layer of the same chain
                        const int rm0 = (tile / p.num_n_tiles) * kTileM + m_rank_off;
                        wait_blocks_ready(p.wait_flags, p.wait_expected, max(0, (rm0 - p.wait_margin) >> 7),
                                          min(p.wait_blocks - 1, (rm0 + kBlockM - 1 + p.wait_margin) >> 7), dep_ready);
                    }
                    const int col = half < 0 ? pf_chunk * 32 : half_chunk_col<BN>(pf_chunk, half);
                    mbar_expect_tx(&res_full[pf_buf], kChunkBytes);
                    tma_load_2d(ring + pf_buf * kChunkBytes, &tmap_res, &res_full[pf_buf],
                                p.res_choff + (tile % p.num_n_tiles) * BN + col, (tile / p.num_n_tiles) * kTileM + m_rank_off);
                    if ((pf_chunk += c_step) >= (half < 0 ? kChunks : kChunks / 2)) { pf_chunk = c_first; pf_item += tile_step; }
                }
                if (++pf_buf == nb) pf_buf = 0;
            };
            // OUT_PHASE by TMA (OutDesc::tma == 2): the 128 staged rows are pixels m0 .. m0+127 of the padded domain (m0 and the row
            // pitch are even: 64 pairs of the 5-D phase view).  Per image row the tile touches: a run that reaches the row's end is
            // ONE 64-pair box (TMA clips the overrun); a run that ends with the tile is covered by two overlapping boxes of the
            // largest power of two that fits (same data written twice where they overlap).  Box starts are never negative.
            const int dom_h = p.dom_plane / p.dom_w, half_w = p.dom_w >> 1;
            auto phase_store = [&](const uint8_t* sbuf, const OutDesc& od, int c0, int m0) {
                const CUtensorMap* maps = reinterpret_cast<const CUtensorMap*>(od.aux);
                const int pairs = (min(m0 + kBlockM, p.m_total) - m0) >> 1;
                const int R = (int)__umul64hi((unsigned long long)m0, p.magic_w);    // image row (over all images) of pixel m0
                int img = (int)__umul64hi((unsigned long long)m0, p.magic_plane);
                int y = R - img * dom_h;
                int a = (m0 - R * p.dom_w) >> 1;                                     // first pair of the run inside its row
                const int nmax2 = 2 * od.nmax;
                for (int q = 0; q < pairs;) {
                    const int n = min(half_w - a, pairs - q);
                    const int yh = y >> 1, j = (y & 1) * nmax2 + img;
                    if (a + n == half_w) {
                        tma_store_5d(sbuf + q * 128, maps + 6, c0, 0, a, yh, j);
                    } else {
                        const int k = 31 - __clz(n), sz = 1 << k;
                        tma_store_5d(sbuf + q * 128, maps + k, c0, 0, a, yh, j);
                        if (n > sz) tma_store_5d(sbuf + (q + n - sz) * 128, maps + k, c0, 0, a + n - sz, yh, j);
                    }
                    q += n; a = 0;
                    if (++y == dom_h) { y = 0; ++img; }
                }
            };
            if (has_res)
                for (int i = 0; i < nb; ++i) prefetch_res();      // every buffer starts free
            int buf = 0, prev = -1; uint32_t sph = 0;
            for (int item = tile_first; item < num_items; item += tile_step) {
                int tile, half;
                decode_item(item, split_from, tile, half);
                const int n_chunks = half < 0 ? kChunks : kChunks / 2;
                const int m0 = (tile / p.num_n_tiles) * kTileM + m_rank_off;
                const int n0 = (tile % p.num_n_tiles) * BN;
#pragma unroll 1
                for (int c = c_first; c < n_chunks; c += c_step) {
                    const int col = half < 0 ? c * 32 : half_chunk_col<BN>(c, half);
                    mbar_wait(&staged[buf], sph);
                    if (any_tma) {
                        const uint8_t* sbuf = ring + buf * kChunkBytes;
                        if (p.out[0].tma == 1) tma_store_2d(sbuf, &tmap_out0, p.out[0].choff + n0 + col, m0);
                        else if (p.out[0].tma == 2) phase_store(sbuf, p.out[0], p.out[0].choff + n0 + col, m0);
                        if (p.out[1].tma == 1) tma_store_2d(sbuf, &tmap_out1, p.out[1].choff + n0 + col, m0);
                        else if (p.out[1].tma == 2) phase_store(sbuf, p.out[1], p.out[1].choff + n0 + col, m0);
                        bulk_commit();
                    }
                    if (prev >= 0) {
                        if (any_tma) bulk_wait_read(1);             // the previous chunk's store has read its buffer
                        if (has_res) prefetch_res(); else mbar_arrive(&buf_free[prev]);
                    }
                    prev = buf;
                    if (++buf == nb) { buf = 0; sph ^= 1; }
                }
                if (p.sig_flags != nullptr) {
                    // this tile's stores have been performed: publish its row block (async-proxy writes -> generic release).
                    // The store warp has slack in the layers that signal (few, long tiles per CTA).
                    bulk_wait_complete(0);
                    fence_proxy_async_all();
                    red_release_gpu_add(p.sig_flags + (m0 >> 7), 1);
                }
            }
            if (prev >= 0) {
                if (any_tma) bulk_wait_read(0);
                if (has_res) prefetch_res(); else mbar_arrive(&buf_free[prev]);
            }
            if (any_tma) bulk_wait_all();                           // shared memory must outlive the last TMA store
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (CTA2) cluster_sync_all();     // neither CTA frees TMEM / exits while its peer may still signal or read
    if (warp == 1) {
        tc_fence_after();
        if constexpr (CTA2) tmem_dealloc_pair(tmem_base, kTmemCols);
        else tmem_dealloc(tmem_base, kTmemCols);
    }
    if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 32 + 22] = globaltimer_ns();
}

}  // namespace fvy
