/* ORACLE — test infrastructure only.  CPU restatement (plain C) of the reference's
 * post-processing: anchor decode, letterbox correction, IoU and greedy NMS.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * link or call this file.  The product path never does.
 *
 * Every function cites the reference lines (under /root/reference/) it restates.  Data are
 * structure-of-arrays instead of the reference's list of BoundBox objects; candidate ORDER is the
 * reference's (cell row-major, anchor ascending).
 *
 * Arithmetic modes:
 *   arith = 0  "f64": scalar math in double — the promotion rules of the reference's pinned
 *                     environment (NumPy 1.x: python-int (+) np.float32 -> float64).
 *   arith = 1  "f32": scalar math in float — NumPy >= 2 (NEP 50) rules, i.e. what the reference's
 *                     own code does when it is executed in this container.
 * In both modes sigmoid/exp are float32 array ops in the reference (yolov3_detect.py:343-344,
 * 375-376).  exp() is restated as the correctly rounded float32 result, (float)exp((double)x);
 * the add and divide of the sigmoid are IEEE float32 operations.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/Makefile).  No FMA contraction.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float orc_expf_cr(float x) { return (float)exp((double)x); }

/* yolov3_detect.py:180-181  _sigmoid on a float32 array: 1. / (1. + np.exp(-x)) */
float orc_sigmoid(float x) {
    volatile float e = orc_expf_cr(-x);
    volatile float s = 1.0f + e;
    return 1.0f / s;
}

/* yolov3_detect.py:165-178  _interval_overlap on integers */
static inline int64_t orc_overlap_i(int64_t x1, int64_t x2, int64_t x3, int64_t x4) {
    if (x3 < x1) {
        if (x4 < x1) return 0;
        return (x2 < x4 ? x2 : x4) - x1;
    }
    if (x2 < x3) return 0;
    return (x2 < x4 ? x2 : x4) - x3;
}

/* yolov3_detect.py:183-194  bbox_iou on integer boxes {xmin,ymin,xmax,ymax}.
 * Returns float(intersect)/union as an IEEE double divide; union==0 gives NaN (numpy-int path
 * of FaceDetector.detect: nan + RuntimeWarning; python-int path raises ZeroDivisionError). */
double orc_bbox_iou_i(const int64_t* a, const int64_t* b) {
    int64_t iw = orc_overlap_i(a[0], a[2], b[0], b[2]);
    int64_t ih = orc_overlap_i(a[1], a[3], b[1], b[3]);
    int64_t inter = iw * ih;
    int64_t w1 = a[2] - a[0], h1 = a[3] - a[1];
    int64_t w2 = b[2] - b[0], h2 = b[3] - b[1];
    int64_t uni = w1 * h1 + w2 * h2 - inter;
    if (uni == 0) return NAN;
    return (double)inter / (double)uni;
}

/* yolov3_detect.py:335-387  decode_netout for ONE scale of ONE image.
 *   netout: (gh, gw, 3*(5+nb_class)) float32 raw logits (not modified)
 *   anchors6: the scale's 3 (w,h) pairs; anchor_mask3 bit b set => anchor b is decoded
 *             (reference mask, :354-362: scale0 -> 0b010, scale1 -> 0b101, scale2 -> 0b010)
 * Outputs (appended from index 0, at most cap): box[n][4] = xmin,ymin,xmax,ymax normalised;
 * objness[n]; classes[n][nb_class]; cell[n] = (row*gw+col)*3+b.  Returns n (or -1 if cap hit). */
int orc_decode_netout(const float* netout, int gh, int gw, int nb_class, const int* anchors6,
                      int anchor_mask3, double obj_thresh, int net_h, int net_w, int arith,
                      double* box, float* objness, float* classes, int* cell, int cap) {
    const int ch = 5 + nb_class;
    int n = 0;
    for (int i = 0; i < gh * gw; ++i) {
        int row = i / gw, col = i % gw;                       /* :349-350 */
        for (int b = 0; b < 3; ++b) {
            if (!((anchor_mask3 >> b) & 1)) continue;         /* :354-362 */
            const float* t = netout + ((size_t)i * 3 + b) * ch;
            float obj = orc_sigmoid(t[4]);                    /* :344 */
            if (arith == 0 ? ((double)obj < obj_thresh) : (obj < (float)obj_thresh)) continue;   /* :368 */
            if (n >= cap) return -1;
            float sx = orc_sigmoid(t[0]), sy = orc_sigmoid(t[1]);   /* :343 */
            float ew = orc_expf_cr(t[2]), eh = orc_expf_cr(t[3]);   /* :375-376 np.exp on f32 */
            double x, y, w, h, x0, y0, x1, y1;
            if (arith == 0) {
                x = ((double)col + (double)sx) / (double)gw;   /* :373 */
                y = ((double)row + (double)sy) / (double)gh;   /* :374 */
                w = (double)anchors6[2 * b + 0] * (double)ew / (double)net_w;   /* :375 */
                h = (double)anchors6[2 * b + 1] * (double)eh / (double)net_h;   /* :376 */
                x0 = x - w / 2; y0 = y - h / 2; x1 = x + w / 2; y1 = y + h / 2;  /* :383 */
            } else {
                volatile float fx = (float)col + sx; fx = fx / (float)gw;
                volatile float fy = (float)row + sy; fy = fy / (float)gh;
                volatile float fw = (float)anchors6[2 * b + 0] * ew; fw = fw / (float)net_w;
                volatile float fh = (float)anchors6[2 * b + 1] * eh; fh = fh / (float)net_h;
                volatile float hw = fw / 2.0f, hh = fh / 2.0f;
                volatile float a0 = fx - hw, a1 = fy - hh, a2 = fx + hw, a3 = fy + hh;
                x0 = a0; y0 = a1; x1 = a2; y1 = a3;
            }
            box[4 * n + 0] = x0; box[4 * n + 1] = y0; box[4 * n + 2] = x1; box[4 * n + 3] = y1;
            objness[n] = obj;
            for (int c = 0; c < nb_class; ++c) classes[(size_t)n * nb_class + c] = orc_sigmoid(t[5 + c]);  /* :344 */
            cell[n] = i * 3 + b;
            ++n;
        }
    }
    return n;
}

/* yolov3_detect.py:389-404  correct_yolo_boxes (and _v2, :406-424, same arithmetic).
 * Reproduces the `new_h = net_w` else-branch (:394), truncation toward zero, no clamping. */
void orc_correct_yolo_boxes(const double* box, int n, int image_h, int image_w, int net_h, int net_w,
                            int arith, int64_t* ibox) {
    double new_w, new_h;
    if ((double)net_w / image_w < (double)net_h / image_h) {   /* :390 */
        new_w = net_w;
        new_h = ((double)image_h * net_w) / image_w;
    } else {
        new_h = net_w;                                          /* :394 (sic) */
        new_w = ((double)image_w * net_h) / image_h;
    }
    double x_offset = (net_w - new_w) / 2. / net_w, x_scale = new_w / net_w;   /* :398 */
    double y_offset = (net_h - new_h) / 2. / net_h, y_scale = new_h / net_h;   /* :399 */
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < 4; ++k) {
            double off = (k & 1) ? y_offset : x_offset, sc = (k & 1) ? y_scale : x_scale;
            int dim = (k & 1) ? image_h : image_w;
            if (arith == 0) {
                double v = (box[4 * i + k] - off) / sc * dim;   /* :401-404 */
                ibox[4 * i + k] = (int64_t)v;
            } else {
                volatile float v = (float)box[4 * i + k] - (float)off;
                v = v / (float)sc;
                v = v * (float)dim;
                ibox[4 * i + k] = (int64_t)v;
            }
        }
    }
}

typedef struct { float key; int idx; } orc_sortrec;
static int orc_cmp_desc(const void* pa, const void* pb) {
    const orc_sortrec* a = (const orc_sortrec*)pa; const orc_sortrec* b = (const orc_sortrec*)pb;
    if (a->key > b->key) return -1;
    if (a->key < b->key) return 1;
    return (a->idx > b->idx) - (a->idx < b->idx);   /* tie rule of the new framework: index ascending */
}

/* yolov3_detect.py:426-444  do_nms  (do_nms_v2, :446-458, is the nb_class==1 case).
 *   ibox[n][4] integer boxes; classes[n][nb_class] scores, suppressed entries set to 0 in place.
 * Sort = np.argsort(-score) (:433) with ties broken by candidate index ascending (the reference's
 * default introsort leaves tie order undefined). */
void orc_do_nms(const int64_t* ibox, int n, int nb_class, double nms_thresh, float* classes) {
    if (n <= 0) return;
    orc_sortrec* ord = (orc_sortrec*)malloc(sizeof(orc_sortrec) * (size_t)n);
    for (int c = 0; c < nb_class; ++c) {
        for (int i = 0; i < n; ++i) { ord[i].key = classes[(size_t)i * nb_class + c]; ord[i].idx = i; }
        qsort(ord, (size_t)n, sizeof(orc_sortrec), orc_cmp_desc);
        for (int i = 0; i < n; ++i) {
            int ii = ord[i].idx;
            if (classes[(size_t)ii * nb_class + c] == 0) continue;         /* :438 */
            for (int j = i + 1; j < n; ++j) {
                int jj = ord[j].idx;
                double iou = orc_bbox_iou_i(ibox + 4 * (size_t)ii, ibox + 4 * (size_t)jj);
                if (iou >= nms_thresh) classes[(size_t)jj * nb_class + c] = 0;   /* :443-444 (NaN => false) */
            }
        }
    }
    free(ord);
}

/* face_detection.py:899-932  the decode half of FaceDetector.detect for one (S/cell)^2 x 6 map.
 *   cands: (grid, grid, 6) float32 raw linear outputs (not modified); cell_px = image_size // 13.
 * Outputs in (row, col) order: ibox[n][4] (np.int64 in the reference), objness[n], score[n],
 * cell[n] = row*grid+col.  Returns n. */
int orc_fd6_decode(const float* cands, int grid, int image_size, int cell_px, double face_conf_th,
                   int arith, int64_t* ibox, float* objness, float* score, int* cell) {
    int n = 0;
    for (int i = 0; i < grid; ++i) {
        for (int j = 0; j < grid; ++j) {
            const float* t = cands + ((size_t)i * grid + j) * 6;
            float obj = orc_sigmoid(t[0]);                    /* :904 */
            volatile float sc = obj * orc_sigmoid(t[5]);      /* :905 float32 product */
            int pass_th = arith == 0 ? ((double)sc >= face_conf_th) : (sc >= (float)face_conf_th);
            if (!(obj > 0.f && pass_th)) continue;            /* :909 */
            double bx = t[1] > 0.f ? (double)t[1] : 0., by = t[2] > 0.f ? (double)t[2] : 0.;   /* :912-913 */
            double bw = t[3] > 0.f ? (double)t[3] : 0., bh = t[4] > 0.f ? (double)t[4] : 0.;   /* :914-915 */
            /* NaN regressors: np.max([nan,0.]) is nan and int(nan) raises in the reference; here they clamp to 0. */
            int64_t px = (int64_t)(bx * cell_px); if (px > cell_px - 1) px = cell_px - 1; px += (int64_t)cell_px * j;  /* :919 */
            int64_t py = (int64_t)(by * cell_px); if (py > cell_px - 1) py = cell_px - 1; py += (int64_t)cell_px * i;  /* :920 */
            double pw = bw * image_size; if (pw > image_size) pw = image_size;   /* :921 */
            double ph = bh * image_size; if (ph > image_size) ph = image_size;   /* :922 */
            int64_t hw = (int64_t)(pw / 2), hh = (int64_t)(ph / 2);
            int64_t xmin = px - hw; if (xmin < 0) xmin = 0;                       /* :925 */
            int64_t ymin = py - hh; if (ymin < 0) ymin = 0;                       /* :926 */
            int64_t xmax = px + hw; if (xmax > image_size - 1) xmax = image_size - 1;   /* :927 */
            int64_t ymax = py + hh; if (ymax > image_size - 1) ymax = image_size - 1;   /* :928 */
            ibox[4 * n + 0] = xmin; ibox[4 * n + 1] = ymin; ibox[4 * n + 2] = xmax; ibox[4 * n + 3] = ymax;
            objness[n] = obj; score[n] = sc; cell[n] = i * grid + j;
            ++n;
        }
    }
    return n;
}

/* face_detection.py:942-947  survivors with get_score() > 0, ascending by score (ties: index
 * ascending), first num_cands.  get_score clips to 1.0 (yolov3_detect.py:151-155).
 * Writes candidate indices to out; returns how many. */
int orc_fd6_select(const float* score, int n, int num_cands, int* out) {
    orc_sortrec* ord = (orc_sortrec*)malloc(sizeof(orc_sortrec) * (size_t)(n > 0 ? n : 1));
    int m = 0;
    for (int i = 0; i < n; ++i) {
        float s = score[i] < 1.0f ? score[i] : 1.0f;
        if (s > 0.f) { ord[m].key = -s; ord[m].idx = i; ++m; }
    }
    qsort(ord, (size_t)m, sizeof(orc_sortrec), orc_cmp_desc);   /* desc on -s == asc on s */
    int k = m < num_cands ? m : num_cands;
    for (int i = 0; i < k; ++i) out[i] = ord[i].idx;
    free(ord);
    return k;
}
