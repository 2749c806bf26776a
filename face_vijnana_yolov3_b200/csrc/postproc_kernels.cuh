// Post-processing kernels: anchor decode (+ letterbox correction), deterministic score sort,
// bitmask-IoU, greedy sweep, detection assembly.  HBM-bound integer / byte work.
//
// Reference semantics restated on the device (paths under /root/reference/src/space/):
//   decode_netout            yolov3_detect.py:335-387      correct_yolo_boxes   :389-404
//   _interval_overlap        :165-178                      bbox_iou             :183-194
//   do_nms / do_nms_v2       :426-458                      BoundBox.get_score   :151-155
//   FaceDetector.detect      face_detection.py:899-947
//
// Exactness rules: every float/double operation that the reference performs as a separate
// rounded operation is an explicit IEEE intrinsic (__fadd_rn, __ddiv_rn, ...) so that no FMA
// contraction can change a result; exp() is evaluated in double and rounded once to float.
// IoU is int64 arithmetic followed by one correctly rounded double divide, exactly the
// reference's `float(intersect) / union`.
#pragma once
#include <climits>

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fvy {

constexpr int kCoordLimit = (1 << 30) - 1;

// ---------------------------------------------------------------- small helpers
__device__ __forceinline__ float exp_cr_f32(float x) { return (float)exp((double)x); }
// _sigmoid on a float32 array: 1. / (1. + np.exp(-x))   (yolov3_detect.py:180-181)
__device__ __forceinline__ float sigmoid_ref(float x) {
    const float e = exp_cr_f32(-x);
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, e));
}

// Exclusive block scan of a 0/1 flag; returns this thread's offset, total through *total (same for all threads).
// blockDim.x must be a multiple of 32 and <= 1024.  `ws` = 33 ints of shared memory.
__device__ __forceinline__ int block_scan_flag(bool flag, int* ws, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    const int within = __popc(bal & ((1u << lane) - 1u));
    __syncthreads();                    // protect ws from the previous call
    if (lane == 0) ws[warp] = __popc(bal);
    __syncthreads();
    if (warp == 0) {
        int v = lane < nwarps ? ws[lane] : 0;
        int incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        ws[lane] = incl - v;
        if (lane == 31) ws[32] = incl;
    }
    __syncthreads();
    *total = ws[32];
    return ws[warp] + within;
}

__device__ __forceinline__ int trunc_to_i32(double v, int* range_flag) {
    // Python int(): truncation toward zero, unbounded.  Here: |v| must stay below 2^30.
    if (!(v > -(double)kCoordLimit && v < (double)kCoordLimit)) { *range_flag = 1; return v < 0 ? -kCoordLimit : kCoordLimit; }
    return (int)v;
}

// ---------------------------------------------------------------- yolo3 decode
struct DecodeArgs {
    const float* out[3];     // (B, gh, gw, C) fp32 logits per scale
    int gh[3], gw[3];
    int nb_class;
    int anchors[18];
    unsigned anchor_mask;    // bit 3*scale + b
    double obj_thresh;
    int net_h, net_w;
    int arith;               // 0 = f64, 1 = f32
    const int* image_hw;     // [B][2] or nullptr
    int cap;                 // slot size per image
    double* nbox;            // [B][cap][4] or nullptr
    int* ibox;               // [B][cap][4] or nullptr
    float* objness;          // [B][cap] or nullptr
    float* classes;          // [B][cap][nb_class] or nullptr
    int* cand;               // [B][cap] or nullptr
    int* counts;             // [B]
    int* status;             // [0] |= 1 range error
};

struct LetterboxConst { double x_off, x_scale, y_off, y_scale; int image_h, image_w; };

// correct_yolo_boxes prologue (yolov3_detect.py:390-399), incl. the `new_h = net_w` else-branch.
__device__ __forceinline__ LetterboxConst letterbox_const(int image_h, int image_w, int net_h, int net_w) {
    double new_w, new_h;
    if (__ddiv_rn((double)net_w, (double)image_w) < __ddiv_rn((double)net_h, (double)image_h)) {
        new_w = (double)net_w;
        new_h = __ddiv_rn((double)((long long)image_h * net_w), (double)image_w);
    } else {
        new_h = (double)net_w;
        new_w = __ddiv_rn((double)((long long)image_w * net_h), (double)image_h);
    }
    LetterboxConst c;
    c.x_off = __ddiv_rn(__ddiv_rn(__dsub_rn((double)net_w, new_w), 2.0), (double)net_w);
    c.x_scale = __ddiv_rn(new_w, (double)net_w);
    c.y_off = __ddiv_rn(__ddiv_rn(__dsub_rn((double)net_h, new_h), 2.0), (double)net_h);
    c.y_scale = __ddiv_rn(new_h, (double)net_h);
    c.image_h = image_h; c.image_w = image_w;
    return c;
}

// int((v - offset) / scale * image_dim)   (yolov3_detect.py:401-404)
__device__ __forceinline__ int correct_coord(double v, double off, double scale, int dim, int arith, int* range_flag) {
    if (arith == 0) {
        const double t = __dmul_rn(__ddiv_rn(__dsub_rn(v, off), scale), (double)dim);
        return trunc_to_i32(t, range_flag);
    }
    const float t = __fmul_rn(__fdiv_rn(__fsub_rn((float)v, (float)off), (float)scale), (float)dim);
    return trunc_to_i32((double)t, range_flag);
}

// Grid (chunks, images): a block of 256 threads decodes kDecodeChunk consecutive candidate slots of one image (4 per thread, in the
// reference's (scale, row, col, anchor) order), scans its pass flags and writes compacted records - so candidate order is the
// reference's.  The position of a chunk's first survivor is the sum of the survivor counts of the image's earlier chunks: chunks take
// a ticket when they START (so every earlier chunk of the image is resident or finished: no deadlock whatever else shares the GPU),
// publish their count with a release store and read the earlier ones with acquire loads ("chained scan").  The six floats of a
// face-head record (24 bytes, 8-byte aligned) are read as three float2.  `sync` = [images][1 + kDecodeMaxChunks] ints, zeroed by the
// caller: [0] ticket, [1 + c] = (count << 1) | 1 once chunk c has published.
constexpr int kDecodeThreads = 256, kDecodePer = 4, kDecodeChunk = kDecodeThreads * kDecodePer, kDecodeMaxChunks = 63;
__device__ __forceinline__ void st_release_gpu(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_gpu_i(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__global__ void __launch_bounds__(kDecodeThreads) decode_yolo_kernel(const DecodeArgs a, int* __restrict__ sync) {
    __shared__ int ws[33];
    __shared__ int s_chunk, s_prefix;
    __shared__ LetterboxConst lb;
    const int img = blockIdx.y;
    const int ch = 5 + a.nb_class;
    const int C = 3 * ch;
    const int n0 = 3 * a.gh[0] * a.gw[0], n1 = n0 + 3 * a.gh[1] * a.gw[1], n2 = n1 + 3 * a.gh[2] * a.gw[2];
    int* isync = sync + (size_t)img * (1 + kDecodeMaxChunks);
    if (threadIdx.x == 0) {
        s_chunk = atomicAdd(isync, 1);
        if (a.image_hw != nullptr) lb = letterbox_const(a.image_hw[2 * img], a.image_hw[2 * img + 1], a.net_h, a.net_w);
    }
    __syncthreads();
    const int chunk = s_chunk;
    const int g0 = chunk * kDecodeChunk + threadIdx.x * kDecodePer;
    // pass flags of this thread's four slots (objectness only: one exp per masked-in slot)
    const float* rec[kDecodePer];
    float obj[kDecodePer];
    int sc[kDecodePer], cell[kDecodePer], bb[kDecodePer];
    unsigned pass = 0;
#pragma unroll
    for (int k = 0; k < kDecodePer; ++k) {
        const int g = g0 + k;
        rec[k] = nullptr; obj[k] = 0.f; sc[k] = 0; cell[k] = 0; bb[k] = 0;
        if (g < n2) {
            const int s = g < n0 ? 0 : (g < n1 ? 1 : 2);
            const int local = g - (s == 0 ? 0 : (s == 1 ? n0 : n1));
            const int c = local / 3, b = local - 3 * c;
            if ((a.anchor_mask >> (3 * s + b)) & 1u) {                                  // :354-362
                const float* t = a.out[s] + ((size_t)img * a.gh[s] * a.gw[s] + c) * C + b * ch;
                const float o = sigmoid_ref(__ldg(t + 4));                              // :344
                const bool below = a.arith == 0 ? ((double)o < a.obj_thresh) : (o < (float)a.obj_thresh);
                if (!below) { pass |= 1u << k; rec[k] = t; obj[k] = o; sc[k] = s; cell[k] = c; bb[k] = b; }   // :368
            }
        }
    }
    // exclusive scan of the per-thread survivor counts inside the chunk
    const int mine = __popc(pass);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int v = lane < kDecodeThreads / 32 ? ws[lane] : 0;
        int in2 = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, in2, d);
            if (lane >= d) in2 += t;
        }
        ws[lane] = in2 - v;
        if (lane == 31) ws[32] = in2;
        // publish this chunk's count (lane 31 holds the block total), then add up the earlier chunks' (spinning until each has published)
        if (lane == 31) st_release_gpu(isync + 1 + chunk, (in2 << 1) | 1);
        int before = 0;
        for (int c = lane; c < chunk; c += 32) {
            int v2;
            for (unsigned spin = 0; ((v2 = ld_acquire_gpu_i(isync + 1 + c)) & 1) == 0; ++spin) {
                __nanosleep(40);
                if (spin > (1u << 24)) { printf("fvy: decode chunk wait timed out (image %d chunk %d of %d)\n", img, c, chunk); __trap(); }
            }
            before += v2 >> 1;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) before += __shfl_xor_sync(0xffffffffu, before, d);
        if (lane == 0) s_prefix = before;
    }
    __syncthreads();
    const int total = ws[32];
    int pos = s_prefix + ws[warp] + (incl - mine);
    if (threadIdx.x == 0 && (chunk + 1) * kDecodeChunk >= n2) a.counts[img] = s_prefix + total;      // the image's last chunk
    int range_flag = 0;
#pragma unroll
    for (int k = 0; k < kDecodePer; ++k) {
        if (!((pass >> k) & 1u)) continue;
        const int my = pos++;
        if (my >= a.cap) continue;
        const float* t = rec[k];
        const int s = sc[k], b = bb[k];
        float r0, r1, r2, r3;
        if (ch == 6) {                   // 24-byte record, 8-byte aligned: three 8-byte loads
            const float2 p0 = __ldg(reinterpret_cast<const float2*>(t)), p1 = __ldg(reinterpret_cast<const float2*>(t) + 1);
            r0 = p0.x; r1 = p0.y; r2 = p1.x; r3 = p1.y;
        } else { r0 = __ldg(t); r1 = __ldg(t + 1); r2 = __ldg(t + 2); r3 = __ldg(t + 3); }
        const int gw = a.gw[s], gh = a.gh[s];
        const int row = cell[k] / gw, col = cell[k] - row * gw;                     // :349-350
        const float sx = sigmoid_ref(r0), sy = sigmoid_ref(r1);                     // :343
        const float ew = exp_cr_f32(r2), eh = exp_cr_f32(r3);                       // :375-376
        const int aw = a.anchors[6 * s + 2 * b], ah = a.anchors[6 * s + 2 * b + 1];
        double x0, y0, x1, y1;
        if (a.arith == 0) {
            const double x = __ddiv_rn(__dadd_rn((double)col, (double)sx), (double)gw);   // :373
            const double y = __ddiv_rn(__dadd_rn((double)row, (double)sy), (double)gh);   // :374
            const double w = __ddiv_rn(__dmul_rn((double)aw, (double)ew), (double)a.net_w);   // :375
            const double h = __ddiv_rn(__dmul_rn((double)ah, (double)eh), (double)a.net_h);   // :376
            const double hw = __ddiv_rn(w, 2.0), hh = __ddiv_rn(h, 2.0);
            x0 = __dsub_rn(x, hw); y0 = __dsub_rn(y, hh); x1 = __dadd_rn(x, hw); y1 = __dadd_rn(y, hh);   // :383
        } else {
            const float x = __fdiv_rn(__fadd_rn((float)col, sx), (float)gw);
            const float y = __fdiv_rn(__fadd_rn((float)row, sy), (float)gh);
            const float w = __fdiv_rn(__fmul_rn((float)aw, ew), (float)a.net_w);
            const float h = __fdiv_rn(__fmul_rn((float)ah, eh), (float)a.net_h);
            const float hw = __fdiv_rn(w, 2.0f), hh = __fdiv_rn(h, 2.0f);
            x0 = __fsub_rn(x, hw); y0 = __fsub_rn(y, hh); x1 = __fadd_rn(x, hw); y1 = __fadd_rn(y, hh);
        }
        const size_t o = (size_t)img * a.cap + my;
        if (a.nbox) { double* d = a.nbox + 4 * o; d[0] = x0; d[1] = y0; d[2] = x1; d[3] = y1; }
        if (a.ibox && a.image_hw) {
            int4 q;
            q.x = correct_coord(x0, lb.x_off, lb.x_scale, lb.image_w, a.arith, &range_flag);
            q.y = correct_coord(y0, lb.y_off, lb.y_scale, lb.image_h, a.arith, &range_flag);
            q.z = correct_coord(x1, lb.x_off, lb.x_scale, lb.image_w, a.arith, &range_flag);
            q.w = correct_coord(y1, lb.y_off, lb.y_scale, lb.image_h, a.arith, &range_flag);
            reinterpret_cast<int4*>(a.ibox)[o] = q;
        }
        if (a.objness) a.objness[o] = obj[k];
        if (a.classes)
            for (int c = 0; c < a.nb_class; ++c) a.classes[o * a.nb_class + c] = sigmoid_ref(__ldg(t + 5 + c));   // :344
        if (a.cand) a.cand[o] = g0 + k;
    }
    if (range_flag) atomicOr(a.status, 1);
}

__global__ void correct_boxes_kernel(const double* nbox, int n, int image_h, int image_w, int net_h, int net_w, int arith,
                                     int* ibox, int* status) {
    const LetterboxConst lb = letterbox_const(image_h, image_w, net_h, net_w);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int rf = 0;
    int4 q;
    q.x = correct_coord(nbox[4 * i + 0], lb.x_off, lb.x_scale, image_w, arith, &rf);
    q.y = correct_coord(nbox[4 * i + 1], lb.y_off, lb.y_scale, image_h, arith, &rf);
    q.z = correct_coord(nbox[4 * i + 2], lb.x_off, lb.x_scale, image_w, arith, &rf);
    q.w = correct_coord(nbox[4 * i + 3], lb.y_off, lb.y_scale, image_h, arith, &rf);
    reinterpret_cast<int4*>(ibox)[i] = q;
    if (rf) atomicOr(status, 1);
}

// ---------------------------------------------------------------- fd6 decode (FaceDetector.detect, face_detection.py:900-932)
struct DecodeFd6Args {
    const float* cands;   // (B, g, g, 6) raw linear outputs
    int grid;             // g = net/32
    int image_size;       // nn_arch.image_size
    int cell_px;          // image_size // 13  (face_detection.py:325)
    double face_conf_th;
    int arith;
    int cap;
    int* ibox; float* objness; float* score; int* cand; int* counts;
};

__global__ void __launch_bounds__(512) decode_fd6_kernel(const DecodeFd6Args a) {
    __shared__ int ws[33];
    const int img = blockIdx.x;
    const int ncell = a.grid * a.grid;
    int base = 0;
    for (int start = 0; start < ncell; start += blockDim.x) {
        const int g = start + threadIdx.x;
        bool pass = false;
        float obj = 0.f, sc = 0.f;
        const float* t = nullptr;
        if (g < ncell) {
            t = a.cands + ((size_t)img * ncell + g) * 6;
            obj = sigmoid_ref(__ldg(t + 0));                               // :904
            sc = __fmul_rn(obj, sigmoid_ref(__ldg(t + 5)));                // :905
            const bool ge = a.arith == 0 ? ((double)sc >= a.face_conf_th) : (sc >= (float)a.face_conf_th);
            pass = obj > 0.f && ge;                                        // :909
        }
        int total;
        const int pos = base + block_scan_flag(pass, ws, &total);
        if (pass && pos < a.cap) {
            const int i = g / a.grid, j = g - i * a.grid;
            const float r1 = __ldg(t + 1), r2 = __ldg(t + 2), r3 = __ldg(t + 3), r4 = __ldg(t + 4);
            const double bx = r1 > 0.f ? (double)r1 : 0.0, by = r2 > 0.f ? (double)r2 : 0.0;   // :912-913
            const double bw = r3 > 0.f ? (double)r3 : 0.0, bh = r4 > 0.f ? (double)r4 : 0.0;   // :914-915
            const double lim = (double)kCoordLimit;
            double tx = __dmul_rn(bx, (double)a.cell_px); tx = tx < lim ? tx : lim;
            double ty = __dmul_rn(by, (double)a.cell_px); ty = ty < lim ? ty : lim;
            int px = (int)tx; px = (px < a.cell_px - 1 ? px : a.cell_px - 1) + a.cell_px * j;   // :919
            int py = (int)ty; py = (py < a.cell_px - 1 ? py : a.cell_px - 1) + a.cell_px * i;   // :920
            double pw = __dmul_rn(bw, (double)a.image_size); pw = pw < (double)a.image_size ? pw : (double)a.image_size;   // :921
            double ph = __dmul_rn(bh, (double)a.image_size); ph = ph < (double)a.image_size ? ph : (double)a.image_size;   // :922
            const int hw = (int)__ddiv_rn(pw, 2.0), hh = (int)__ddiv_rn(ph, 2.0);
            int4 q;
            q.x = max(px - hw, 0); q.y = max(py - hh, 0);                                       // :925-926
            q.z = min(px + hw, a.image_size - 1); q.w = min(py + hh, a.image_size - 1);         // :927-928
            const size_t o = (size_t)img * a.cap + pos;
            reinterpret_cast<int4*>(a.ibox)[o] = q;
            a.objness[o] = obj; a.score[o] = sc; a.cand[o] = g;
        }
        base += total;
    }
    if (threadIdx.x == 0) a.counts[img] = base;
}

// ---------------------------------------------------------------- IoU
// _interval_overlap (yolov3_detect.py:165-178)
__device__ __forceinline__ long long interval_overlap(long long x1, long long x2, long long x3, long long x4) {
    if (x3 < x1) {
        if (x4 < x1) return 0;
        return (x2 < x4 ? x2 : x4) - x1;
    }
    if (x2 < x3) return 0;
    return (x2 < x4 ? x2 : x4) - x3;
}
struct IouParts { long long inter, uni; };
__device__ __forceinline__ IouParts iou_parts(const int4 a, const int4 b) {
    const long long iw = interval_overlap(a.x, a.z, b.x, b.z);
    const long long ih = interval_overlap(a.y, a.w, b.y, b.w);
    IouParts r;
    r.inter = iw * ih;                                                     // :187
    const long long w1 = (long long)a.z - a.x, h1 = (long long)a.w - a.y;
    const long long w2 = (long long)b.z - b.x, h2 = (long long)b.w - b.y;
    r.uni = w1 * h1 + w2 * h2 - r.inter;                                   // :192
    return r;
}
// bbox_iou(...) >= thresh  (yolov3_detect.py:194, 443).  union == 0 -> nan -> false.
__device__ __forceinline__ bool iou_ge(const int4 a, const int4 b, double th, bool zero_ge_th) {
    const IouParts p = iou_parts(a, b);
    if (p.uni == 0) return false;
    if (p.inter == 0) return zero_ge_th;
    return __ddiv_rn((double)p.inter, (double)p.uni) >= th;
}
// Same decision for WELL-FORMED boxes (xmin <= xmax, ymin <= ymax) with precomputed areas and th > 0, arranged so that
// the common case costs a handful of 32-bit instructions:
//   * no strict overlap  => _interval_overlap returns 0 on one axis => intersect == 0 => iou is 0 (or nan) => not >= th;
//   * otherwise intersect = iw*ih exactly in int64, union = area_a + area_b - intersect, and
//     fl(intersect/union) >= th is decided by comparing intersect with th*union in double when the two differ by more
//     than 2 ulp (fl is monotone), and by the correctly rounded divide itself inside that band.
//   * before any 64-bit work a float estimate decides every pair that is not within 1 % of the threshold (the estimate's own
//     error is ~1e-6 relative): fp64 and int64 run at a small fraction of the fp32 rate on this GPU.
__device__ __forceinline__ bool iou_ge_fast(const int4 a, long long area_a, float area_a_f, const int4 b, long long area_b, float area_b_f,
                                            double th, float th_lo, float th_hi) {
    if (a.z <= b.x || b.z <= a.x || a.w <= b.y || b.w <= a.y) return false;
    const int iw = min(a.z, b.z) - max(a.x, b.x);
    const int ih = min(a.w, b.w) - max(a.y, b.y);
    {
        const float xf = (float)iw * (float)ih;
        const float uf = area_a_f + area_b_f - xf;
        if (uf > 0.f && area_a_f < 1e30f && area_b_f < 1e30f) {
            if (xf < th_lo * uf) return false;
            if (xf > th_hi * uf) return true;
        }
    }
    const long long inter = (long long)iw * (long long)ih;
    const long long uni = area_a + area_b - inter;
    if (uni == 0) return false;
    const double x = (double)inter, u = (double)uni;
    const double t = __dmul_rn(th, u);
    if (x > __dmul_rn(t, 1.0 + 0x1p-51)) return true;
    if (x < __dmul_rn(t, 1.0 - 0x1p-51)) return false;
    return __ddiv_rn(x, u) >= th;
}
// The same pre-filter on float copies of the boxes (no int -> float conversions, one fused disjointness test) for tiles whose
// coordinates are all below 2^23 in magnitude: every coordinate, every width / height and hence every decision is exactly the
// one iou_ge_fast takes.  0 = below the band, 1 = above, 2 = inside the 1 % band (the caller runs the exact int64 / fp64 test).
__device__ __forceinline__ int iou_prefilter_f(const float4 a, float area_a_f, const float4 b, float area_b_f, float th_lo, float th_hi) {
    const float iw = fminf(a.z, b.z) - fmaxf(a.x, b.x);
    const float ih = fminf(a.w, b.w) - fmaxf(a.y, b.y);
    if (iw <= 0.f || ih <= 0.f) return 0;           // disjoint or touching (well-formed boxes): IoU 0 < th
    const float xf = iw * ih;
    const float uf = area_a_f + area_b_f - xf;
    if (uf > 0.f && area_a_f < 1e30f && area_b_f < 1e30f) {
        if (xf < th_lo * uf) return 0;
        if (xf > th_hi * uf) return 1;
    }
    return 2;
}
__global__ void bbox_iou_kernel(const int4* a, const int4* b, int n, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const IouParts p = iou_parts(a[i], b[i]);
    out[i] = p.uni == 0 ? __longlong_as_double(0x7ff8000000000000LL) : __ddiv_rn((double)p.inter, (double)p.uni);
}

// ---------------------------------------------------------------- IoU of FLOAT boxes
// bbox_iou / do_nms are type-generic in the reference (yolov3_detect.py:165-194, 426-444): called before correct_yolo_boxes, or from
// evaluate.py:69,275, the BoundBox coordinates are floats and every operation is a float operation.  Two arithmetic modes, as for the
// decode: FVY_ARITH_F64 = Python floats / np.float64 (every step a double operation, `float(intersect) / union` a double divide),
// FVY_ARITH_F32 = np.float32 coordinates under NumPy >= 2 (every step a float operation; the Python float `float(intersect)` is a
// weak scalar, so the divide is a float32 divide, and `>= nms_thresh` compares in float32).  Each operation is separately rounded.
__device__ __forceinline__ double fp_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float fp_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double fp_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float fp_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double fp_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float fp_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double fp_div(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float fp_div(float a, float b) { return __fdiv_rn(a, b); }
template <typename T>
__device__ __forceinline__ T interval_overlap_fp(T x1, T x2, T x3, T x4) {      // :165-178; Python min(a, b) = b if b < a else a
    if (x3 < x1) {
        if (x4 < x1) return (T)0;
        return fp_sub(x4 < x2 ? x4 : x2, x1);
    }
    if (x2 < x3) return (T)0;
    return fp_sub(x4 < x2 ? x4 : x2, x3);
}
template <typename T>
__device__ __forceinline__ T iou_fp(const T* a, const T* b) {                    // :183-194; union == 0 -> nan / inf like NumPy
    const T iw = interval_overlap_fp(a[0], a[2], b[0], b[2]);
    const T ih = interval_overlap_fp(a[1], a[3], b[1], b[3]);
    const T inter = fp_mul(iw, ih);
    const T w1 = fp_sub(a[2], a[0]), h1 = fp_sub(a[3], a[1]), w2 = fp_sub(b[2], b[0]), h2 = fp_sub(b[3], b[1]);
    const T uni = fp_sub(fp_add(fp_mul(w1, h1), fp_mul(w2, h2)), inter);
    return fp_div(inter, uni);
}
__global__ void bbox_iou_fp_kernel(const double* a, const double* b, int n, int arith, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (arith == 0) out[i] = iou_fp<double>(a + 4 * i, b + 4 * i);
    else {
        const float fa[4] = {(float)a[4 * i], (float)a[4 * i + 1], (float)a[4 * i + 2], (float)a[4 * i + 3]};
        const float fb[4] = {(float)b[4 * i], (float)b[4 * i + 1], (float)b[4 * i + 2], (float)b[4 * i + 3]};
        out[i] = (double)iou_fp<float>(fa, fb);
    }
}

// ---------------------------------------------------------------- deterministic sort
__device__ __forceinline__ uint32_t float_orderable(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // ascending uint order == ascending float order
}

// Bitonic sort of `np2` 64-bit keys (np2 = power of two >= 64) held in shared or global memory, by a block of whole warps.
// Every warp owns aligned 64-key segments; the compare-exchange steps with partner distance j <= 32 stay inside a segment, so a
// whole run of them (the tail j = 32 .. 1 of every stage, and all of the stages k <= 64) needs only __syncwarp; the block-wide
// barrier is paid for the steps with j >= 64 alone: 21 instead of 78 barriers at np2 = 4096.
// compare-exchange of pair p of a step with partner distance j (a power of two): elements i and i + j, i = (p / j) * 2 j + p % j -
// every thread that calls this does work (indexing by element instead leaves half of the lanes idle in every step)
__device__ __forceinline__ void bitonic_pair(unsigned long long* keys, int p, int j, int k) {
    const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
    const unsigned long long a = keys[i], b = keys[i + j];
    const bool up = (i & k) == 0;
    if ((a > b) == up) { keys[i] = b; keys[i + j] = a; }
}
__device__ void bitonic_sort_u64(unsigned long long* keys, int np2) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    // stages k = 2 .. 64: entirely inside 64-key segments (32 pairs per step: one per lane)
    for (int seg = warp * 64; seg < np2; seg += nwarps * 64) {
        for (int k = 2; k <= 64; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                bitonic_pair(keys, (seg >> 1) + lane, j, k);
                __syncwarp();
            }
    }
    __syncthreads();
    for (int k = 128; k <= np2; k <<= 1) {
        for (int j = k >> 1; j >= 64; j >>= 1) {
            for (int p = threadIdx.x; p < (np2 >> 1); p += blockDim.x) bitonic_pair(keys, p, j, k);
            __syncthreads();
        }
        for (int seg = warp * 64; seg < np2; seg += nwarps * 64) {
            for (int j = 32; j > 0; j >>= 1) {
                bitonic_pair(keys, (seg >> 1) + lane, j, k);
                __syncwarp();
            }
        }
        __syncthreads();
    }
}

struct SortArgs {
    const int* ibox;        // [.. ][4]
    const float* classes;   // [..][nb_class]
    const int* counts;      // [B]
    int seg_stride;         // entries between image segments in ibox/classes
    int nb_class, cls;      // class being processed
    int capP;               // scratch stride per image (multiple of 64)
    int descending;         // 1: score desc (do_nms), 0: score asc (detect's final argsort)
    int* order;             // [B][capP] sorted candidate indices
    int4* sbox;             // [B][capP] boxes in sorted order (may be nullptr)
    uint4* srow;            // [B][capP] half-precision record of each sorted box for nms_mask_kernel's pre-filter (with sbox; may be nullptr)
    unsigned char* sflag;   // [B][capP/32] per 32 sorted boxes: bit 0 = all well-formed, bit 1 = all have a half record
    unsigned long long* rowflag;   // [B][capP/64] zeroed here for the mask kernel (may be nullptr)
    unsigned long long* gkeys;   // [B][np2max] global scratch when the keys do not fit in shared memory
    int smem_keys;          // number of keys that fit in dynamic shared memory
    int np2max;
};

constexpr int kHalfCoord = 60000;            // |coordinate| bound for a half-precision record (finite after directed rounding)
constexpr long long kHalfArea = 16000000;    // area bound: area / 256 stays below the largest half

// One block per image: order = argsort(-score) with ties by index ascending (yolov3_detect.py:433, 447).
__global__ void __launch_bounds__(1024) sort_scores_kernel(const SortArgs a) {
    extern __shared__ unsigned long long skeys[];
    const int img = blockIdx.x;
    const int n = min(a.counts[img], a.seg_stride);
    if (a.rowflag)      // consumed by the kernels behind this one
        for (int i = threadIdx.x; i < (a.capP >> 6); i += blockDim.x) a.rowflag[(size_t)img * (a.capP >> 6) + i] = 0ull;
    if (n <= 0) return;
    int np2 = 64;
    while (np2 < n) np2 <<= 1;
    unsigned long long* keys = np2 <= a.smem_keys ? skeys : a.gkeys + (size_t)img * a.np2max;
    const size_t seg = (size_t)img * a.seg_stride;
    for (int i = threadIdx.x; i < np2; i += blockDim.x) {
        unsigned long long k = ~0ull;
        if (i < n) {
            uint32_t o = float_orderable(a.classes[(seg + i) * a.nb_class + a.cls]);
            if (a.descending) o = ~o;
            k = ((unsigned long long)o << 32) | (uint32_t)i;
        }
        keys[i] = k;
    }
    __syncthreads();
    bitonic_sort_u64(keys, np2);
    const bool recs = a.sbox != nullptr && a.srow != nullptr;
    const int n64 = (n + 63) & ~63;            // whole 64-blocks: both flag bytes of the last block are written
    for (int i = threadIdx.x; i < n64; i += blockDim.x) {
        bool wf = true, hk = true;
        if (i < n) {
            const int idx = (int)(keys[i] & 0xffffffffu);
            a.order[(size_t)img * a.capP + i] = idx;
            if (a.sbox) {
                const int4 b = reinterpret_cast<const int4*>(a.ibox)[seg + idx];
                a.sbox[(size_t)img * a.capP + i] = b;
                if (recs) {
                    uint4 rec = make_uint4(0u, 0u, 0u, 0u);
                    wf = b.z >= b.x && b.w >= b.y;
                    const long long area = ((long long)b.z - b.x) * ((long long)b.w - b.y);
                    hk = (unsigned)(b.x + kHalfCoord) <= 2u * kHalfCoord && (unsigned)(b.y + kHalfCoord) <= 2u * kHalfCoord &&
                         (unsigned)(b.z + kHalfCoord) <= 2u * kHalfCoord && (unsigned)(b.w + kHalfCoord) <= 2u * kHalfCoord &&
                         area >= 0 && area < kHalfArea;
                    if (hk) {
                        // the box grown to half precision (minima rounded down, maxima up: an overlap of the exact boxes is an
                        // overlap of these) and its area / 256 rounded to nearest (relative error 2^-11, see nms_mask_kernel)
                        const unsigned x = __half_as_ushort(__float2half_rd((float)b.x)), y = __half_as_ushort(__float2half_rd((float)b.y));
                        const unsigned z = __half_as_ushort(__float2half_ru((float)b.z)), w = __half_as_ushort(__float2half_ru((float)b.w));
                        const unsigned ar = __half_as_ushort(__float2half_rn((float)area * (1.0f / 256.0f)));
                        rec = make_uint4(x | (y << 16), z | (w << 16), ar, __float_as_uint((float)area));    // .w: the float area of phase 2
                    }
                    a.srow[(size_t)img * a.capP + i] = rec;
                }
            }
        }
        if (recs) {     // i < n64 is warp-uniform (n64 is a multiple of 64)
            const unsigned bw = __ballot_sync(0xffffffffu, wf), bk = __ballot_sync(0xffffffffu, hk);
            if ((threadIdx.x & 31) == 0) a.sflag[(size_t)img * (a.capP >> 5) + (i >> 5)] = (unsigned char)((bw == 0xffffffffu ? 1 : 0) | (bk == 0xffffffffu ? 2 : 0));
        }
    }
}

// ---------------------------------------------------------------- bitmask IoU
__device__ __forceinline__ __half2 u32_as_half2(unsigned v) { return *reinterpret_cast<const __half2*>(&v); }
struct MaskArgs {
    const int4* sbox;       // [B][capP]
    const int* counts;      // [B]
    int seg_stride;
    int batch, capP, words; // words = capP / 64
    double th;
    const uint4* srow;      // [B][capP] half records of the sorted boxes (sort_scores_kernel)
    const unsigned char* sflag;   // [B][capP/32]
    unsigned long long* mask;   // [B][capP][words]; only words >= row/64 are written
    unsigned long long* rowflag; // [B][words], zeroed by the caller; see SweepArgs
};

// Persistent blocks of 64 threads; one 64x64 tile per iteration: thread t owns sorted row r*64+t and
// produces the 64-bit word of column block c (bit j set <=> IoU(row, c*64+j) >= th and c*64+j > row).  Work items are the
// upper-triangle tiles only, dealt round-robin: every block gets the same number of real tiles.
//
// Fast path (every box of the row block and of the column block well-formed and with a half record - sort_scores_kernel's flags -
// and th > 0), two phases:
//  1. a packed-half pre-filter over all 64 columns, two columns per instruction (HSET2 masks): a bit survives when the boxes,
//     grown to half precision, overlap AND the areas are close enough for the IoU to reach the threshold at all
//     (IoU <= min / max of the areas; boxes of different anchors - areas 480 .. 30 888 px^2 - cannot however much they overlap).
//     It is a SUPERSET of the pairs with IoU >= th: growing a box keeps every strict overlap, and with a = rn(A / 256), b likewise
//     (relative error e = 2^-11 each) and t = half(0.99 th), A >= th B implies a >= A (1 - e) / 256 >= th B (1 - e) / 256 >=
//     0.99 th B (1 + e)^3 / 256 >= rn(t b): the 1 % of slack dwarfs the three roundings.  12 instructions per two columns.
//     The staged columns are ordered so that shifting the running word left once per step leaves bit j = column j.
//  2. the surviving bits (a few per cent at the usual candidate densities) take the float pre-filter (iou_prefilter_f) on float
//     copies of the column boxes staged in shared memory, and the 1 % band around the threshold the exact int64 / fp64 test on the
//     integer boxes: the decisions of the reference's bbox_iou, bit for bit.
// Every other tile (malformed boxes, coordinates beyond +-60 000, th <= 0) takes the plain per-pair loops.
__global__ void __launch_bounds__(64, 16) nms_mask_kernel(const MaskArgs a) {
    __shared__ uint4 s_cols[64];          // per step: {x2, y2, z2, w2} and {b2, t*b2, -, -}: two columns per half2
    __shared__ float4 s_cf[64];           // float copies of the column boxes and of their areas for phase 2
    __shared__ float s_caf[64];
    __shared__ int tile_prefix[1025];     // batch <= 1024
    for (int b = threadIdx.x; b < a.batch; b += blockDim.x) {
        const int n = min(a.counts[b], a.seg_stride);
        const int nb = (n + 63) >> 6;
        tile_prefix[b + 1] = nb * (nb + 1) / 2;          // upper-triangle tiles only (c >= r): every work item is a real tile
    }
    if (threadIdx.x == 0) tile_prefix[0] = 0;
    __syncthreads();
    if (threadIdx.x == 0)
        for (int b = 0; b < a.batch; ++b) tile_prefix[b + 1] += tile_prefix[b];
    __syncthreads();
    const int total = tile_prefix[a.batch];
    const bool zero_ge = 0.0 >= a.th;
    const float th_lo = (float)(a.th * 0.99), th_hi = (float)(a.th * 1.01);     // 1 % either side of the threshold
    const __half2 th2 = __float2half2_rn(th_lo);
    int img = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
        while (tile_prefix[img + 1] <= t) ++img;          // t is increasing
        const int n = min(a.counts[img], a.seg_stride);
        const int nb = (n + 63) >> 6;
        const int lt = t - tile_prefix[img];
        // row r holds the nb - r tiles c = r .. nb - 1; rows 0 .. r-1 hold S(r) = r nb - r (r - 1) / 2 tiles: largest r with S(r) <= lt
        const float d = (float)(2 * nb + 1);
        int r = (int)((d - sqrtf(fmaxf(d * d - 8.0f * (float)lt, 0.f))) * 0.5f);
        r = max(0, min(r, nb - 1));
        while (r + 1 < nb && (r + 1) * nb - ((r + 1) * r >> 1) <= lt) ++r;
        while (r > 0 && r * nb - (r * (r - 1) >> 1) > lt) --r;
        const int c = r + lt - (r * nb - (r * (r - 1) >> 1));
        const int4* sb = a.sbox + (size_t)img * a.capP;
        const uint4* sr = a.srow + (size_t)img * a.capP;
        const unsigned char* fl = a.sflag + (size_t)img * (a.capP >> 5);
        const unsigned flags = (unsigned)fl[2 * r] & fl[2 * r + 1] & fl[2 * c] & fl[2 * c + 1];
        const bool wellformed = (flags & 1u) != 0;
        const bool fast = flags == 3u && !zero_ge;
        const int col = c * 64 + threadIdx.x;
        const int row = r * 64 + threadIdx.x;
        __syncthreads();                                   // the previous tile's readers are done with s_cols
        if (fast) {
            const uint4 rec = col < n ? sr[col] : make_uint4(0u, 0u, 0u, 0u);
            const int4 cb = col < n ? sb[col] : make_int4(0, 0, 0, 0);
            s_cf[threadIdx.x] = make_float4((float)cb.x, (float)cb.y, (float)cb.z, (float)cb.w);
            s_caf[threadIdx.x] = __uint_as_float(rec.w);
            // step `it` of word w handles columns 32 w + 15 - it (low halves) and 32 w + 31 - it (high halves)
            const int j = threadIdx.x;
            unsigned short* e = reinterpret_cast<unsigned short*>(s_cols) + (size_t)((j >> 5) * 16 + 15 - (j & 15)) * 16 + ((j >> 4) & 1);
            const __half bh = __ushort_as_half((unsigned short)(rec.z & 0xffffu));
            e[0] = (unsigned short)(rec.x & 0xffffu); e[2] = (unsigned short)(rec.x >> 16);
            e[4] = (unsigned short)(rec.y & 0xffffu); e[6] = (unsigned short)(rec.y >> 16);
            e[8] = (unsigned short)(rec.z & 0xffffu); e[10] = __half_as_ushort(__hmul(__low2half(th2), bh));
        }
        __syncthreads();
        if (row < n) {
            unsigned long long word = 0;
            const int jmax = min(64, n - c * 64);
            const int j0 = (c == r ? threadIdx.x + 1 : 0);
            if (fast) {
                const uint4 rec = sr[row];
                const __half2 mex = u32_as_half2(__byte_perm(rec.x, 0u, 0x1010)), mey = u32_as_half2(__byte_perm(rec.x, 0u, 0x3232));
                const __half2 mez = u32_as_half2(__byte_perm(rec.y, 0u, 0x1010)), mew = u32_as_half2(__byte_perm(rec.y, 0u, 0x3232));
                const __half2 mya = u32_as_half2(__byte_perm(rec.z, 0u, 0x1010));
                const __half2 myta = __hmul2(th2, mya);
                unsigned acc[2];
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                    unsigned v = 0;
#pragma unroll
                    for (int it = 0; it < 16; ++it) {
                        const uint4 q = s_cols[(w * 16 + it) * 2];
                        const uint2 ar = *reinterpret_cast<const uint2*>(&s_cols[(w * 16 + it) * 2 + 1]);
                        const unsigned m1 = __hgt2_mask(mez, u32_as_half2(q.x)), m2 = __hgt2_mask(u32_as_half2(q.z), mex);
                        const unsigned m3 = __hgt2_mask(mew, u32_as_half2(q.y)), m4 = __hgt2_mask(u32_as_half2(q.w), mey);
                        const unsigned m5 = __hge2_mask(mya, u32_as_half2(ar.y)), m6 = __hge2_mask(u32_as_half2(ar.x), myta);
                        v = (v << 1) + ((m1 & m2 & m3) & (m4 & m5 & m6) & 0x00010001u);
                    }
                    acc[w] = v;
                }
                const unsigned long long upto = jmax >= 64 ? ~0ull : ((1ull << jmax) - 1ull);
                const unsigned long long from = j0 >= 64 ? 0ull : ~((1ull << j0) - 1ull);
                const unsigned long long ov = (((unsigned long long)acc[1] << 32) | acc[0]) & upto & from;
                if (ov) {
                    const int4 me = sb[row];
                    const long long my_area = ((long long)me.z - me.x) * ((long long)me.w - me.y);
                    const float my_area_f = (float)my_area;
                    const float4 mef = make_float4((float)me.x, (float)me.y, (float)me.z, (float)me.w);
                    for (unsigned long long todo = ov; todo; todo &= todo - 1) {
                        const int j = __ffsll((long long)todo) - 1;
                        const float area_b_f = s_caf[j];
                        const int dcs = iou_prefilter_f(mef, my_area_f, s_cf[j], area_b_f, th_lo, th_hi);
                        if (dcs == 1) word |= 1ull << j;
                        else if (dcs == 2) {          // inside the 1 % band: the exact test on the integer boxes
                            const int4 b = sb[c * 64 + j];
                            if (iou_ge_fast(me, my_area, my_area_f, b, ((long long)b.z - b.x) * ((long long)b.w - b.y), area_b_f, a.th, th_lo, th_hi)) word |= 1ull << j;
                        }
                    }
                }
            } else {
                const int4 me = sb[row];
                if (wellformed && !zero_ge) {
                    const long long my_area = ((long long)me.z - me.x) * ((long long)me.w - me.y);
                    const float my_area_f = (float)my_area;
                    for (int j = j0; j < jmax; ++j) {
                        const int4 b = sb[c * 64 + j];
                        const long long area_b = ((long long)b.z - b.x) * ((long long)b.w - b.y);
                        if (iou_ge_fast(me, my_area, my_area_f, b, area_b, (float)area_b, a.th, th_lo, th_hi)) word |= 1ull << j;
                    }
                } else {
                    for (int j = j0; j < jmax; ++j)
                        if (iou_ge(me, sb[c * 64 + j], a.th, zero_ge)) word |= 1ull << j;
                }
            }
            a.mask[((size_t)img * a.capP + row) * a.words + c] = word;
            if (c > r && word) atomicOr(&a.rowflag[(size_t)img * a.words + r], 1ull << threadIdx.x);
        }
    }
}

// The same bitmask for FLOAT boxes (fvy_nms_fp): boxes [..][4] double in candidate order, read through the sorted order; one
// thread per sorted row, plain loop over the 64 columns of a tile (this path serves host-side callers, not the hot path).
struct MaskFpArgs {
    const double* box;      // [..][4], segment b at b * seg_stride
    const int* order;       // [B][capP]
    const int* counts;
    int seg_stride, batch, capP, words, arith;
    double th;
    unsigned long long* mask; unsigned long long* rowflag;
};
__global__ void __launch_bounds__(64) nms_mask_fp_kernel(const MaskFpArgs a) {
    __shared__ double cbox[64][4];
    for (int img = 0; img < a.batch; ++img) {
        const int n = min(a.counts[img], a.seg_stride);
        const int nb = (n + 63) >> 6;
        const double* bx = a.box + (size_t)img * a.seg_stride * 4;
        const int* ord = a.order + (size_t)img * a.capP;
        for (int t = blockIdx.x; t < nb * nb; t += gridDim.x) {
            const int r = t / nb, c = t - r * nb;
            if (c < r) continue;
            const int col = c * 64 + threadIdx.x, row = r * 64 + threadIdx.x;
            __syncthreads();
            if (col < n) { const double* q = bx + 4 * (size_t)ord[col]; cbox[threadIdx.x][0] = q[0]; cbox[threadIdx.x][1] = q[1]; cbox[threadIdx.x][2] = q[2]; cbox[threadIdx.x][3] = q[3]; }
            __syncthreads();
            if (row >= n) continue;
            const double* q = bx + 4 * (size_t)ord[row];
            const double me[4] = {q[0], q[1], q[2], q[3]};
            const float mef[4] = {(float)q[0], (float)q[1], (float)q[2], (float)q[3]};
            const float thf = (float)a.th;
            unsigned long long word = 0;
            const int jmax = min(64, n - c * 64);
            for (int j = (c == r ? threadIdx.x + 1 : 0); j < jmax; ++j) {
                bool ge;
                if (a.arith == 0) ge = iou_fp<double>(me, cbox[j]) >= a.th;
                else {
                    const float cf[4] = {(float)cbox[j][0], (float)cbox[j][1], (float)cbox[j][2], (float)cbox[j][3]};
                    ge = iou_fp<float>(mef, cf) >= thf;
                }
                if (ge) word |= 1ull << j;
            }
            a.mask[((size_t)img * a.capP + row) * a.words + c] = word;
            if (c > r && word) atomicOr(&a.rowflag[(size_t)img * a.words + r], 1ull << threadIdx.x);
        }
    }
}

// In-place sigmoid of one scale's netout as decode_netout applies it to the CALLER's array (yolov3_detect.py:343-344):
// channels 0, 1 and 4.. of every (cell, anchor); tw, th stay raw.
__global__ void netout_sigmoid_kernel(float* netout, long long n_boxes, int ch) {
    const long long total = n_boxes * ch;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % ch);
        if (c == 2 || c == 3) continue;
        netout[i] = sigmoid_ref(netout[i]);
    }
}

// ---------------------------------------------------------------- cal_mAP_fd matching (evaluate.py:41-100)
// Per image: IoU of every (ground-truth face i, detection j) pair with bbox_iou on the CSV's numbers (float64 arithmetic: pandas
// hands np.int64 / np.float64 to BoundBox; integer-valued inputs are exact in double), pairs with IoU > 0 kept (:69-70), then the
// reference's greedy assignment: repeatedly take the pair with the largest IoU, write it to detection j, drop every pair of
// row i and of column j (:84-96).  Ties (equal IoU; the reference's sort is unstable): smaller (i, j) first.
// One block per image; the pair matrix lives in global scratch (pair_off[img] .. pair_off[img + 1]).
struct MapMatchArgs {
    const double* gt;        // [n_gt][4]  x1, y1, x2, y2
    const double* det;       // [n_det][4]
    const int* gt_off;       // [n_img + 1]
    const int* det_off;      // [n_img + 1]
    const long long* pair_off;   // [n_img + 1]
    double* pair_iou;        // scratch [pair_off[n_img]]
    double* det_iou;         // [n_det] out: the IoU assigned to the detection, or -1 (the reference's initial value, :31)
    int* img_any;            // [n_img] out: 1 if the image has at least one pair with IoU > 0 (else the reference skips it, :76)
};
__global__ void __launch_bounds__(256) map_match_kernel(const MapMatchArgs a) {
    __shared__ double s_best[8];
    __shared__ long long s_idx[8];
    __shared__ long long s_pick;
    const int img = blockIdx.x;
    const int g0 = a.gt_off[img], G = a.gt_off[img + 1] - g0;
    const int d0 = a.det_off[img], D = a.det_off[img + 1] - d0;
    double* m = a.pair_iou + a.pair_off[img];
    const long long n = (long long)G * D;
    for (int j = threadIdx.x; j < D; j += blockDim.x) a.det_iou[d0 + j] = -1.0;
    int any = 0;
    for (long long p = threadIdx.x; p < n; p += blockDim.x) {
        const int i = (int)(p / D), j = (int)(p - (long long)i * D);
        const double v = iou_fp<double>(a.gt + 4 * (size_t)(g0 + i), a.det + 4 * (size_t)(d0 + j));
        const bool pos = v > 0.0;              // nan (zero union) and 0 are dropped
        m[p] = pos ? v : -1.0;
        any |= pos ? 1 : 0;
    }
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) a.img_any[img] = any;
    if (!any) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (;;) {
        double best = -1.0; long long bi = -1;
        for (long long p = threadIdx.x; p < n; p += blockDim.x) {       // ascending p: the first of equal values wins
            const double v = m[p];
            if (v > best) { best = v; bi = p; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, best, o);
            const long long oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi >= 0 && (bi < 0 || oi < bi))) { best = ov; bi = oi; }
        }
        if (lane == 0) { s_best[warp] = best; s_idx[warp] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double b = s_best[0]; long long k = s_idx[0];
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
                if (s_best[w] > b || (s_best[w] == b && s_idx[w] >= 0 && (k < 0 || s_idx[w] < k))) { b = s_best[w]; k = s_idx[w]; }
            s_pick = b > 0.0 ? k : -1;
            if (b > 0.0) a.det_iou[d0 + (int)(k % D)] = b;                // rel_sol_df.iat[j, -1] = iou   (:92)
        }
        __syncthreads();
        const long long pick = s_pick;
        if (pick < 0) break;
        const int pi = (int)(pick / D), pj = (int)(pick - (long long)pi * D);
        for (long long p = threadIdx.x; p < n; p += blockDim.x) {       // remove assigned samples (:95-96)
            const int i = (int)(p / D), j = (int)(p - (long long)i * D);
            if (i == pi || j == pj) m[p] = -1.0;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- greedy sweep
struct SweepArgs {
    const unsigned long long* mask;
    const int* order;       // [B][capP]
    const int* counts;
    int seg_stride, capP, words;
    int nb_class, cls;
    float* classes;         // in/out
    const unsigned long long* rowflag;   // [B][words]: bit i of word i/64 set <=> sorted row i suppresses something in a later block
};

// One block per image.  Walks the sorted list in blocks of 64.  Per block step:
//   (1) the 64 diagonal words and the "score != 0" bits of the NEXT block are prefetched by warps 2-3 while
//   (2) warp 0 resolves the current block with the words held in registers: only rows that are still candidates AND whose
//       diagonal word hits another candidate can change anything, so the dependent chain visits those rows alone (none at
//       all in the common case of a block without mutual overlaps) instead of all 64;
//   (3) kept rows that suppress something in a LATER block (row flags set by nms_mask_kernel) have their later words
//       OR-ed into the removed mask; rows without such a word cost no memory access.
// Boxes whose score is already 0 never suppress (yolov3_detect.py:438).  Same result as the literal sweep (:440-444).
__global__ void __launch_bounds__(1024) nms_sweep_kernel(const SweepArgs a) {
    extern __shared__ unsigned long long removed[];    // [words]
    __shared__ unsigned long long diag[2][64];
    __shared__ unsigned alive32[2][2];
    __shared__ unsigned long long flagged_w;
    const int img = blockIdx.x;
    const int n = min(a.counts[img], a.seg_stride);
    if (n <= 0) return;
    const int nb = (n + 63) >> 6;
    const size_t seg = (size_t)img * a.seg_stride;
    const int* ord = a.order + (size_t)img * a.capP;
    const unsigned long long* mk = a.mask + (size_t)img * a.capP * a.words;
    const unsigned long long* rf = a.rowflag + (size_t)img * a.words;
    for (int w = threadIdx.x; w < nb; w += blockDim.x) removed[w] = 0;
    auto fetch_block = [&](int blk, int buf, int t /* 0..63 */) {
        const int i = blk * 64 + t;
        unsigned long long d = 0;
        bool alive = false;
        if (i < n) {
            d = mk[(size_t)i * a.words + blk];
            alive = a.classes[(seg + ord[i]) * a.nb_class + a.cls] != 0.f;
        }
        diag[buf][t] = d;
        const unsigned bal = __ballot_sync(0xffffffffu, alive);
        if ((t & 31) == 0) alive32[buf][t >> 5] = bal;
    };
    if (threadIdx.x < 64) fetch_block(0, 0, threadIdx.x);
    for (int blk = 0; blk < nb; ++blk) {
        const int cur = blk & 1;
        __syncthreads();        // diag[cur] / alive32[cur] are in place; every OR into removed[blk] has been made
        if (threadIdx.x >= 64 && threadIdx.x < 128 && blk + 1 < nb) fetch_block(blk + 1, cur ^ 1, threadIdx.x - 64);
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            const unsigned long long d_lo = diag[cur][lane], d_hi = diag[cur][lane + 32];
            const unsigned long long alive = (unsigned long long)alive32[cur][0] | ((unsigned long long)alive32[cur][1] << 32);
            unsigned long long cr = removed[blk];
            const unsigned long long cand = alive & ~cr;
            const bool lo_act = ((cand >> lane) & 1ull) && (d_lo & cand) != 0ull;
            const bool hi_act = ((cand >> (lane + 32)) & 1ull) && (d_hi & cand) != 0ull;
            unsigned long long act = (unsigned long long)__ballot_sync(0xffffffffu, lo_act) |
                                     ((unsigned long long)__ballot_sync(0xffffffffu, hi_act) << 32);
            while (act) {                                   // ascending row order = the reference's score order inside the block
                const int b = __ffsll((long long)act) - 1;
                act &= act - 1;
                const unsigned long long db = __shfl_sync(0xffffffffu, b < 32 ? d_lo : d_hi, b & 31);
                if (!((cr >> b) & 1ull)) cr |= db;
            }
            if (lane == 0) {
                const unsigned long long keep = alive & ~cr;
                removed[blk] = cr; flagged_w = keep & rf[blk];
            }
        }
        __syncthreads();
        const unsigned long long fl = flagged_w;
        const int later = nb - (blk + 1);
        if (fl) {                                           // block-uniform; one pass, every (flagged row, later word) pair in parallel
            for (int idx = threadIdx.x; idx < 64 * later; idx += blockDim.x) {
                const int b = idx / later, w = blk + 1 + (idx - b * later);
                if ((fl >> b) & 1ull) {
                    const unsigned long long m = mk[(size_t)(blk * 64 + b) * a.words + w];
                    if (m) atomicOr(&removed[w], m);
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        if ((removed[i >> 6] >> (i & 63)) & 1ull) a.classes[(seg + ord[i]) * a.nb_class + a.cls] = 0.f;   // :444
}

// The same sweep for segments of up to kSweepMaxBlocks x 64 boxes with every global round trip taken off the per-block chain.
// nms_sweep_kernel pays, per block step, a dependent global read for the next block's "score != 0" bits (order -> classes) and
// another for the later words of the kept rows (~1.9 us per step, 64 us for 2 100 boxes).  Here the diagonal words, the row flags
// and the alive bits of ALL blocks are read up front into shared memory (one parallel pass), and the later words of a block's
// flagged rows are fetched two whole steps before they are needed (two register sets -> a double buffer in shared memory), so a
// step is two barriers and shared-memory work.  Same decisions in the same order: bit-identical result.
constexpr int kSweepMaxBlocks = 96;          // 6 144 boxes: 148 KB of shared memory at the limit
constexpr int kSweepThreads = 512;           // 64 rows x 8 threads: thread t serves row t >> 3, later words (t & 7) + 8 k
constexpr int kSweepPf = (kSweepMaxBlocks - 1 + 7) / 8;      // words a thread may hold in flight per register set
__host__ __device__ inline size_t sweep_small_smem(int nb) { return (size_t)nb * 8 * (3 + 64 + 2 * 64); }
__global__ void __launch_bounds__(kSweepThreads) nms_sweep_small_kernel(const SweepArgs a, int nb_max) {
    extern __shared__ unsigned long long sm[];
    __shared__ unsigned long long flagged_w;
    const int img = blockIdx.x;
    const int n = min(a.counts[img], a.seg_stride);
    if (n <= 0) return;
    const int nb = (n + 63) >> 6;
    unsigned long long* removed = sm;                       // [nb_max]
    unsigned long long* alive = sm + nb_max;                // [nb_max]
    unsigned long long* rfs = sm + 2 * nb_max;              // [nb_max]  row flags (rows with a word in a later block)
    unsigned long long* diag = sm + 3 * nb_max;             // [nb_max][64]
    unsigned long long* pre = diag + (size_t)nb_max * 64;   // [2][64][nb_max]: later words of the flagged rows of a block
    const size_t seg = (size_t)img * a.seg_stride;
    const int* ord = a.order + (size_t)img * a.capP;
    const unsigned long long* mk = a.mask + (size_t)img * a.capP * a.words;
    const unsigned long long* rf = a.rowflag + (size_t)img * a.words;
    const int lane = threadIdx.x & 31;
    for (int w = threadIdx.x; w < nb; w += blockDim.x) { removed[w] = 0; rfs[w] = rf[w]; }
    for (int i0 = (threadIdx.x >> 5) * 32; i0 < nb * 64; i0 += blockDim.x) {       // whole warps: i0 .. i0 + 31 share half a block
        const int i = i0 + lane;
        unsigned long long d = 0;
        bool al = false;
        if (i < n) {
            d = mk[(size_t)i * a.words + (i >> 6)];
            al = a.classes[(seg + ord[i]) * a.nb_class + a.cls] != 0.f;
        }
        diag[i] = d;
        const unsigned bal = __ballot_sync(0xffffffffu, al);
        if (lane == 0) reinterpret_cast<unsigned*>(alive)[i >> 5] = bal;
    }
    __syncthreads();                                         // rfs is read by the prefetches below
    // later words of block c's flagged rows: thread (row b = t >> 3, lane8 = t & 7) holds words w = c + 1 + lane8 + 8 k
    const int prow = threadIdx.x >> 3, pl8 = threadIdx.x & 7;
    unsigned long long pf0[kSweepPf], pf1[kSweepPf];         // blocks of even / odd index: fetched TWO steps before they are stashed
    auto fetch_rows = [&](int c, unsigned long long (&pf)[kSweepPf]) {
        const int later = nb - c - 1;
        const bool on = c < nb && ((rfs[c < nb ? c : 0] >> prow) & 1ull);
        const unsigned long long* src = mk + (size_t)(c * 64 + prow) * a.words + c + 1;
#pragma unroll
        for (int k = 0; k < kSweepPf; ++k) {
            const int w = pl8 + 8 * k;
            pf[k] = (on && w < later) ? src[w] : 0ull;
        }
    };
    auto stash_rows = [&](int c, const unsigned long long (&pf)[kSweepPf]) {
        const int later = nb - c - 1;
        unsigned long long* dst = pre + (size_t)(c & 1) * 64 * nb_max + (size_t)prow * nb_max;
#pragma unroll
        for (int k = 0; k < kSweepPf; ++k) {
            const int w = pl8 + 8 * k;
            if (w < later) dst[w] = pf[k];
        }
    };
    fetch_rows(0, pf0);
    stash_rows(0, pf0);
    fetch_rows(1, pf1);
    fetch_rows(2, pf0);
    auto step = [&](int blk, unsigned long long (&pf)[kSweepPf]) {      // pf: the register set of block blk + 1 (and blk + 3)
        __syncthreads();        // diag / alive / pre[blk & 1] are in place; every OR into removed[blk] has been made
        stash_rows(blk + 1, pf);    // fetched two steps ago; pre[(blk + 1) & 1] was last read in step blk - 1, before the barrier above
        fetch_rows(blk + 3, pf);
        if (threadIdx.x < 32) {
            const unsigned long long d_lo = diag[blk * 64 + lane], d_hi = diag[blk * 64 + lane + 32];
            const unsigned long long al = alive[blk];
            unsigned long long cr = removed[blk];
            const unsigned long long cand = al & ~cr;
            const bool lo_act = ((cand >> lane) & 1ull) && (d_lo & cand) != 0ull;
            const bool hi_act = ((cand >> (lane + 32)) & 1ull) && (d_hi & cand) != 0ull;
            unsigned long long act = (unsigned long long)__ballot_sync(0xffffffffu, lo_act) |
                                     ((unsigned long long)__ballot_sync(0xffffffffu, hi_act) << 32);
            while (act) {                                   // ascending row order = the reference's score order inside the block
                const int b = __ffsll((long long)act) - 1;
                act &= act - 1;
                const unsigned long long db = __shfl_sync(0xffffffffu, b < 32 ? d_lo : d_hi, b & 31);
                if (!((cr >> b) & 1ull)) cr |= db;
            }
            if (lane == 0) {
                const unsigned long long keep = al & ~cr;
                removed[blk] = cr; flagged_w = keep & rfs[blk];
            }
        }
        __syncthreads();
        const unsigned long long fl = flagged_w;
        const int later = nb - (blk + 1);
        if ((fl >> prow) & 1ull) {
            const unsigned long long* src = pre + (size_t)(blk & 1) * 64 * nb_max + (size_t)prow * nb_max;
            for (int w = pl8; w < later; w += 8) {
                const unsigned long long m = src[w];
                if (m) atomicOr(&removed[blk + 1 + w], m);
            }
        }
    };
    for (int blk = 0; blk < nb; blk += 2) {
        step(blk, pf1);
        if (blk + 1 < nb) step(blk + 1, pf0);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        if ((removed[i >> 6] >> (i & 63)) & 1ull) a.classes[(seg + ord[i]) * a.nb_class + a.cls] = 0.f;   // :444
}

// ---------------------------------------------------------------- kept list / detection records
struct FvyDet { int xmin, ymin, xmax, ymax; float objness, score; int label, cand; };

struct AssembleArgs {
    const int* ibox; const float* objness; const float* classes; const int* cand; const int* counts;
    int seg_stride, nb_class;
    int max_out, limit;      // limit = num_cands cap (<= max_out)
    int* kept_idx;           // [B][seg_stride] or nullptr
    int* kept_counts;        // [B] or nullptr
    FvyDet* dets;            // [B][max_out] or nullptr
    int* det_counts;         // [B] or nullptr
};

// One block per image: candidates that keep a class score > 0 after NMS, in candidate order.
__global__ void __launch_bounds__(1024) assemble_yolo_kernel(const AssembleArgs a) {
    __shared__ int ws[33];
    const int img = blockIdx.x;
    const int n = min(a.counts[img], a.seg_stride);
    const size_t seg = (size_t)img * a.seg_stride;
    int base = 0;
    for (int start = 0; start < n; start += blockDim.x) {
        const int i = start + threadIdx.x;
        bool keep = false; float best = 0.f; int label = 0;
        if (i < n) {
            const float* c = a.classes + (seg + i) * a.nb_class;
            best = c[0];
            for (int k = 1; k < a.nb_class; ++k) if (c[k] > best) { best = c[k]; label = k; }   // np.argmax: first maximum
            for (int k = 0; k < a.nb_class; ++k) keep |= c[k] > 0.f;
        }
        int total;
        const int pos = base + block_scan_flag(keep, ws, &total);
        if (keep) {
            if (a.kept_idx) a.kept_idx[seg + pos] = i;
            if (a.dets && pos < a.limit) {
                const int4 q = reinterpret_cast<const int4*>(a.ibox)[seg + i];
                FvyDet d;
                d.xmin = q.x; d.ymin = q.y; d.xmax = q.z; d.ymax = q.w;
                d.objness = a.objness ? a.objness[seg + i] : 0.f;
                d.score = best < 1.0f ? best : 1.0f;                  // get_score clips to 1.0 (:155)
                d.label = label; d.cand = a.cand ? a.cand[seg + i] : i;
                a.dets[(size_t)img * a.max_out + pos] = d;
            }
        }
        base += total;
    }
    if (threadIdx.x == 0) {
        if (a.kept_counts) a.kept_counts[img] = base;
        if (a.det_counts) a.det_counts[img] = min(base, a.limit);
    }
}

// fd6 tail (face_detection.py:942-947): survivors with get_score() > 0, ascending by score (ties: index
// ascending), first num_cands.  `order` = argsort ascending over ALL candidates (sort_scores_kernel with
// descending = 0 after NMS zeroing); zero scores sort first and are skipped.
__global__ void __launch_bounds__(512) assemble_fd6_kernel(const AssembleArgs a, const int* order, int capP) {
    __shared__ int ws[33];
    const int img = blockIdx.x;
    const int n = min(a.counts[img], a.seg_stride);
    const size_t seg = (size_t)img * a.seg_stride;
    int base = 0;
    for (int start = 0; start < n; start += blockDim.x) {
        const int p = start + threadIdx.x;
        bool keep = false; int i = 0; float sc = 0.f;
        if (p < n) {
            i = order[(size_t)img * capP + p];
            sc = a.classes[seg + i];
            sc = sc < 1.0f ? sc : 1.0f;
            keep = sc > 0.f;
        }
        int total;
        const int pos = base + block_scan_flag(keep, ws, &total);
        if (keep && pos < a.limit) {
            const int4 q = reinterpret_cast<const int4*>(a.ibox)[seg + i];
            FvyDet d;
            d.xmin = q.x; d.ymin = q.y; d.xmax = q.z; d.ymax = q.w;
            d.objness = a.objness[seg + i]; d.score = sc; d.label = 0; d.cand = a.cand[seg + i];
            a.dets[(size_t)img * a.max_out + pos] = d;
        }
        base += total;
    }
    if (threadIdx.x == 0) a.det_counts[img] = min(base, a.limit);
}

}  // namespace fvy
