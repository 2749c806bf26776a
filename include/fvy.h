/* fvy.h — C ABI of the B200-native face-detection hot path (libfvy.so).
 *
 * The reference (tonandr/face_vijnana_yolov3) is pure Python with no FFI of its own: the boundary
 * of its hot path is `Model.predict` plus the host functions that follow it.  Each entry point
 * below cites the reference interface it replaces (paths under /root/reference/).  The Python
 * drop-ins in face_vijnana_yolov3_b200/space/ bind these symbols with ctypes.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++ / torch types.
 *   - Every function returns 0 (FVY_OK) or a negative fvy_status; the message is available from
 *     fvy_last_error() (thread-local).  No C++ exception crosses the ABI.
 *   - Data pointers may be HOST or DEVICE pointers (queried with cudaPointerGetAttributes);
 *     host buffers are staged through pinned memory owned by the handle.  The caller owns every
 *     pointer it passes; the handle owns weights, activation arena, tensor maps and scratch.
 *   - A handle is bound to one device and one stream and is not re-entrant; distinct handles are
 *     independent (8 GPUs = 8 handles, one process or thread each).
 *   - There is no CPU fallback: without a CUDA device every compute call fails with FVY_E_CUDA.
 */
#ifndef FVY_H_
#define FVY_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fvy_handle fvy_handle;

typedef enum {
    FVY_OK = 0,
    FVY_E_INVALID = -1,   /* bad argument */
    FVY_E_CUDA = -2,      /* CUDA runtime / driver error, or no device */
    FVY_E_STATE = -3,     /* e.g. forward before load_weights */
    FVY_E_CAPACITY = -4,  /* more candidates than the configured capacity */
    FVY_E_RANGE = -5      /* decoded coordinate outside +-2^30 (reference: unbounded Python int) */
} fvy_status;

enum { FVY_HEAD_YOLO3 = 0, FVY_HEAD_FD6 = 1, FVY_HEAD_NONE = 2 };   /* NONE: post-processing only */
enum { FVY_F32 = 0, FVY_F64 = 1, FVY_U8 = 2 };                      /* dtype of the image tensor */
/* FVY_U8: uint8 pixels 0..255 as imread returns them; the device forms float32(pixel / 255.0 evaluated in float64), i.e. the value
 * `image / 255` (face_detection.py:660, :800) has after Keras casts its input to float32 - same logits, a quarter of the float32
 * bytes on the host-to-device path. */
enum { FVY_ARITH_F64 = 0, FVY_ARITH_F32 = 1 };                     /* decode scalar arithmetic (see fvy_decode) */

/* The fork's anchor mask in decode_netout (src/space/yolov3_detect.py:354-362): bit (3*scale+b). */
#define FVY_ANCHOR_MASK_REFERENCE 0x0AAu
#define FVY_ANCHOR_MASK_ALL 0x1FFu

typedef struct {
    int device;          /* CUDA device ordinal */
    int net_h, net_w;    /* network input size, multiples of 32 (416 / 608) */
    int head;            /* FVY_HEAD_* */
    int nb_class;        /* yolo3 heads: C = 3*(5+nb_class) (yolov3_detect.py:278,294,308 hard-code 80) */
    int bb_info_c_size;  /* fd6 head channels (face_vijnana_yolov3.json:27 -> 6) */
    int max_batch;       /* images per forward call */
    int max_cands;       /* candidate capacity per image for decode/NMS; 0 = every cell x anchor */
    int tile_n_max;      /* 0 = default; upper bound on the GEMM N tile (tuning knob) */
    int flags;           /* 0 = the measured-best schedule, or a combination of FVY_CFG_* (the execution-schedule switches a caller
                            may pin per handle; results are bit-identical either way, tests/test_gpu_parity.py) */
} fvy_config;
/* fvy_config.flags: each bit overrides the corresponding FVY_* environment variable for this handle. */
#define FVY_CFG_NO_GRAPH 0x01u         /* launch the conv stack kernel by kernel instead of replaying a CUDA graph (FVY_GRAPH=0) */
#define FVY_CFG_NO_CHAIN 0x02u         /* no layer chains (conv_chain_kernel), one launch per layer (FVY_CHAIN=0) */
#define FVY_CFG_NO_TILE_FLAGS 0x04u    /* kernel boundaries instead of cross-layer tile dependencies; implies NO_CHAIN (FVY_FLAGS=0) */
#define FVY_CFG_CHAIN_SCHED 0x08u      /* per-pair work lists from the host list schedule inside every chain (FVY_CHAIN_SCHED=1; default: chains of few tile waves) */
#define FVY_CFG_NO_CHAIN_SCHED 0x10u   /* static rotation inside the chains (FVY_CHAIN_SCHED=0) */
#define FVY_CFG_NO_OVERLAP_POST 0x20u  /* asynchronous calls post-process on the main stream (FVY_OVERLAP_POST=0) */
#define FVY_CFG_NO_FUSED_STEM 0x40u    /* conv_0 and conv_1 as two kernels with the 4-phase activation in HBM (FVY_FUSE_STEM=0) */
#define FVY_CFG_NO_COMPACT 0x80u       /* every level stored with a private halo per image, (H+2) x (W+2) (FVY_COMPACT=0) */

/* One detection, 32 bytes.  Mirrors the fields of the reference's BoundBox that survive the hot
 * path (src/space/yolov3_detect.py:126-145): xmin,ymin,xmax,ymax,objness, get_label(), get_score(). */
typedef struct {
    int32_t xmin, ymin, xmax, ymax;
    float objness;
    float score;   /* classes[label] after NMS */
    int32_t label; /* argmax class */
    int32_t cand;  /* candidate index in reference order: yolo3 = offset(scale) + (row*gw+col)*3+b (all-anchor
                      numbering); fd6 = row*13+col */
} fvy_det;

typedef struct {
    double obj_thresh;        /* decode_netout obj_thresh (yolov3_detect.py:335,368) / hps.face_conf_th (face_detection.py:909) */
    double nms_thresh;        /* do_nms nms_thresh (yolov3_detect.py:426,443) / hps.nms_iou_th (face_detection.py:939) */
    int num_cands;            /* fd6: hps.num_cands (face_detection.py:946-947); yolo3: cap on returned boxes, 0 = max_out */
    unsigned anchor_mask;     /* yolo3: FVY_ANCHOR_MASK_* */
    int anchors[18];          /* yolo3: 9 (w,h) pairs, scale 0 (stride 32) first (yolov3_detect.py:558-560) */
    int arith;                /* FVY_ARITH_* */
} fvy_post_params;

const char* fvy_last_error(void);
const char* fvy_version(void);

/* Lifetime.  Replaces model construction: make_yolov3_model() (yolov3_detect.py:217-311) /
 * FaceDetector.__init__ + YOLOV3Base (face_detection.py:312-382, 384-600). */
int fvy_create(const fvy_config* cfg, fvy_handle** out);
void fvy_destroy(fvy_handle* h);

/* Weight ingest.  Replaces WeightReader.load_weights (yolov3_detect.py:90-121): `stream` is the
 * float32 payload of a Darknet .weights file (header stripped) in the order that function reads
 * it; for FVY_HEAD_FD6 the 3x3 head's bias[c] + kernel[c][1024][3][3] follow conv_73.  Folds
 * BatchNorm (eps = 1e-3) in fp32, rounds once to bf16, packs K-major GEMM operands. HOST pointer. */
int fvy_load_weights(fvy_handle* h, const float* stream, size_t n_floats);
/* Number of floats fvy_load_weights expects (= parameter count: 61 576 342 for yolo3, nb_class 1). */
long long fvy_weight_count(const fvy_handle* h);

/* Forward.  Replaces Model.predict (yolov3_detect.py:593, face_detection.py:899).
 *   images: (batch, net_h, net_w, 3) NHWC, dtype FVY_F32 / FVY_F64, values in [0,1] (or FVY_U8, values 0..255).
 *   out0/out1/out2: yolo3 -> (batch, H/32, W/32, C), (batch, H/16, W/16, C), (batch, H/8, W/8, C) fp32;
 *                   fd6   -> out0 = (batch, H/32, W/32, bb_info_c_size), out1 = out2 = NULL.
 *   Any out pointer may be NULL: the logits then stay in the handle for fvy_postprocess. */
int fvy_forward(fvy_handle* h, const void* images, int dtype, int batch, float* out0, float* out1, float* out2);

/* Decode, yolo3.  Replaces decode_netout over the three scales (yolov3_detect.py:335-387, loop at
 * :596-598) and, when image_hw != NULL, correct_yolo_boxes (:389-404).
 *   out0..2: logits as produced by fvy_forward (NULL = use the handle's resident logits).
 *   image_hw: batch x {image_h, image_w} or NULL.
 *   Outputs (each may be NULL), per image b a slot of `cap` = capacity entries:
 *     nbox  [batch][cap][4] double  xmin,ymin,xmax,ymax normalised to the net (BoundBox floats)
 *     ibox  [batch][cap][4] int32   after correct_yolo_boxes (needs image_hw)
 *     objness [batch][cap], classes [batch][cap][nb_class] float, cand [batch][cap] int32
 *     counts [batch] int32
 *   arith: FVY_ARITH_F64 = scalar math in double (NumPy 1.x promotion, the reference's pinned
 *   environment); FVY_ARITH_F32 = float (NumPy >= 2).  sigmoid/exp are float32 array ops in both. */
int fvy_decode(fvy_handle* h, const float* out0, const float* out1, const float* out2, int batch,
               const fvy_post_params* pp, const int* image_hw, int cap, double* nbox, int32_t* ibox, float* objness,
               float* classes, int32_t* cand, int32_t* counts);

/* correct_yolo_boxes / correct_yolo_boxes_v2 on its own (yolov3_detect.py:389-424): n boxes of one image. */
int fvy_correct_boxes(fvy_handle* h, const double* nbox, int n, int image_h, int image_w, int net_h, int net_w,
                      int arith, int32_t* ibox);

/* Greedy per-class NMS.  Replaces do_nms / do_nms_v2 (yolov3_detect.py:426-458) with bbox_iou
 * (:183-194) on integer boxes.  Segment b holds counts[b] boxes starting at entry b*seg_stride.
 *   ibox    [..][4] int32;  classes [..][nb_class] float, IN/OUT: suppressed scores set to 0.
 *   kept_idx (optional) [batch][seg_stride] int32: indices (within the segment, ascending) whose
 *   score for class 0..nb_class-1 is still > 0 for at least one class; kept_counts [batch].
 * Order: score descending, ties by index ascending (the reference's argsort leaves ties undefined). */
int fvy_nms(fvy_handle* h, const int32_t* ibox, const int32_t* counts, int batch, int seg_stride, int nb_class,
            double nms_thresh, float* classes, int32_t* kept_idx, int32_t* kept_counts);

/* Pairwise IoU of integer boxes, float(intersect)/union as an IEEE double (bbox_iou, :183-194);
 * union == 0 -> NaN.  a, b: [n][4] int32, out[n]. */
int fvy_bbox_iou(fvy_handle* h, const int32_t* a, const int32_t* b, int n, double* out);

/* bbox_iou / do_nms on FLOAT boxes.  The reference's functions are type-generic (yolov3_detect.py:165-194, 426-444): called on
 * BoundBox floats - before correct_yolo_boxes, or from evaluate.py:69,275 - every step is a float operation.  Boxes are passed
 * as doubles; arith selects the operation type: FVY_ARITH_F64 = Python floats / np.float64 (double operations, double divide),
 * FVY_ARITH_F32 = np.float32 coordinates under NumPy >= 2 (float operations; `float(intersect) / union` and `>= nms_thresh`
 * are evaluated in float32 because Python scalars are weak there).  union == 0 gives nan / inf as NumPy does (never >= thresh
 * for nan).  fvy_nms_fp: HOST pointers; box [batch][seg_stride][4]; otherwise as fvy_nms. */
int fvy_bbox_iou_fp(fvy_handle* h, const double* a, const double* b, int n, int arith, double* out);
int fvy_nms_fp(fvy_handle* h, const double* box, const int32_t* counts, int batch, int seg_stride, int nb_class,
               double nms_thresh, int arith, float* classes, int32_t* kept_idx, int32_t* kept_counts);
/* The in-place part of decode_netout (yolov3_detect.py:343-344): netout[..., :2] and netout[..., 4:] of one scale's
 * (grid_h, grid_w, 3, 5 + nb_class) array become their sigmoid (1 / (1 + exp(-x)) in float32, exp correctly rounded), in the
 * caller's array (host or device), n_boxes = grid_h * grid_w * 3. */
int fvy_netout_sigmoid(fvy_handle* h, float* netout, long long n_boxes, int nb_class);

/* The matching half of evaluate.cal_mAP_fd (src/space/evaluate.py:41-100), row f-4: for each of n_img images, bbox_iou of every
 * (ground-truth face, detection) pair in float64 (the CSV numbers as pandas hands them to BoundBox), pairs with IoU > 0 kept,
 * then the greedy assignment "largest IoU first, drop its row and column" (:84-96).  gt_box / det_box: [..][4] x1, y1, x2, y2 with
 * the boxes of image k at gt_off[k] .. gt_off[k+1] (det_off likewise); det_iou[n_det]: the IoU assigned to each detection or -1;
 * img_any[n_img]: 1 if the image had any pair with IoU > 0 (the reference skips the others, :76).  Equal IoUs: smaller
 * (face, detection) index first (the reference's sort leaves ties undefined).  HOST pointers. */
int fvy_map_match(fvy_handle* h, const double* gt_box, const int32_t* gt_off, const double* det_box, const int32_t* det_off, int n_img,
                  double* det_iou, int32_t* img_any);

/* Post-processing of resident (or given) logits into detections.
 *   yolo3: decode -> correct_yolo_boxes -> do_nms -> boxes with a surviving class score, candidate order.
 *   fd6:   FaceDetector.detect after predict (face_detection.py:900-947): sigmoid, threshold, box math,
 *          do_nms_v2, score > 0, ascending by score, first num_cands.
 *   dets [batch][max_out], det_counts [batch]. */
int fvy_postprocess(fvy_handle* h, const float* out0, const float* out1, const float* out2, int batch,
                    const fvy_post_params* pp, const int* image_hw, int max_out, fvy_det* dets, int32_t* det_counts);

/* forward + postprocess in one call (yolov3_detect._main_ :593-604 / FaceDetector.detect :899-947). */
int fvy_detect(fvy_handle* h, const void* images, int dtype, int batch, const fvy_post_params* pp, const int* image_hw,
               int max_out, fvy_det* dets, int32_t* det_counts);

/* Introspection for tests / bench: */
int fvy_num_layers(const fvy_handle* h);
/* info[0..11] = idx, cin, cout, k, stride, H_out, W_out, tile_n, tile_k, stages, grid, num_tiles;
 * stages = A slots + 100 * B slots + 10000 * weights resident + 100000 * CTA pair + 1000000 * A slab */
int fvy_layer_info(const fvy_handle* h, int layer, int* info12);
/* Copies layer `layer`'s stored output for `batch` images into dst as dense NHWC float32 (debug / layer-wise parity). */
int fvy_layer_output(fvy_handle* h, int layer, int batch, float* dst_host);
/* Kernels launched by this handle since creation (bench.py reports gpu_launches from it). */
long long fvy_launch_count(const fvy_handle* h);
/* Device-time of the last fvy_forward / fvy_postprocess call in ms (CUDA events on the handle's stream). */
int fvy_last_timing(const fvy_handle* h, float* forward_ms, float* post_ms);
/* Per-layer device time of one forward (synchronises between layers; profiling aid). ms[num_layers]. */
int fvy_profile_layers(fvy_handle* h, int batch, int iters, float* ms);
/* Runs one conv layer (1 warm-up + iters timed launches) on the activations of the last forward; *ms = mean device time. */
int fvy_run_layer(fvy_handle* h, int layer, int batch, int iters, float* ms);
/* Device stopwatch on the handle's stream (CUDA events): start records an event; stop records another,
 * synchronises the stream and returns the elapsed milliseconds.  bench.py brackets its K timed steps with it. */
int fvy_timer_start(fvy_handle* h);
int fvy_timer_stop(fvy_handle* h, float* ms);
/* Mean device time (CUDA events inside the timed region, on the streams the kernels run on) of the forward part and of the
 * post-processing part of the fvy_detect / fvy_detect_async calls made since fvy_timer_start (any number of calls: a ring of 32 event pairs is drained into running sums): what
 * bench.py's roofline entries are computed from.  Synchronises the handle's streams. */
int fvy_timer_breakdown(fvy_handle* h, float* forward_ms_mean, float* post_ms_mean, int* calls);
/* Blocks until all work queued on the handle's streams is done.  Returns the deferred error of an asynchronous call, if any
 * (FVY_E_CAPACITY / FVY_E_RANGE: fvy_detect_async returns before the decode has run). */
int fvy_sync(fvy_handle* h);
/* Async variants for pipelined serving: enqueue only; results valid after fvy_sync. Host buffers must be pinned. */
int fvy_detect_async(fvy_handle* h, const void* images, int dtype, int batch, const fvy_post_params* pp,
                     const int* image_hw, int max_out, fvy_det* dets, int32_t* det_counts);
/* Keras-2.2.4 Adam update over one flat fp32 bucket, all pointers DEVICE memory, enqueued on `cuda_stream` (a cudaStream_t,
 * NULL = default stream):  g' = g * grad_scale;  m = b1 m + (1-b1) g';  v = b2 v + (1-b2) g'^2;  p -= lr_t m / (sqrt(v) + eps).
 * lr_t carries Keras' decay and bias correction (lr / (1 + decay it) * sqrt(1 - b2^t) / (1 - b1^t)).  Replaces the optimizer the
 * reference compiles in src/space/face_detection.py:361-364; grad_scale = 1 / world_size turns the all-reduced SUM of the
 * per-GPU gradients into multi_gpu_model's whole-batch mean (face_detection.py:369). */
int fvy_adam_step(float* param, const float* grad, float* m, float* v, long long n, float lr_t, float beta_1, float beta_2,
                  float epsilon, float grad_scale, void* cuda_stream);
/* Training step, row f-1 (first slice): BatchNormalization in TRAINING mode fused with the LeakyReLU that follows it, forward and
 * backward, over NHWC float32 activations x[rows][C] (rows = batch * H * W, C a multiple of 4), all pointers DEVICE memory, enqueued
 * on `cuda_stream`.  Replaces what Keras runs for every bnorm_i + leaky_i pair of the reference's model when it trains
 * (src/space/yolov3_detect.py:212-213 via src/space/face_detection.py:361-381, :602-630): batch statistics per GPU (as in
 * multi_gpu_model's towers), eps = 1e-3, Keras momentum 0.99 <=> `momentum` = 0.01 here (new = old + momentum (batch - old), the
 * running variance takes the unbiased batch variance).  slope = 0.1 (LeakyReLU(alpha=0.1)); slope = 1 gives a plain BatchNorm.
 *   forward : y = leaky(gamma (x - mean) / sqrt(var + eps) + beta); save_mean / save_invstd [C] are kept for the backward.
 *   backward: dx, dgamma [C], dbeta [C] from dy (gradient w.r.t. y) and the saved statistics.
 * workspace: 2 * C doubles of scratch. */
int fvy_bn_leaky_train_forward(const float* x, long long rows, int C, const float* gamma, const float* beta, float eps, float momentum,
                               float slope, float* running_mean, float* running_var, float* y, float* save_mean, float* save_invstd,
                               double* workspace, void* cuda_stream);
int fvy_bn_leaky_train_backward(const float* x, const float* dy, long long rows, int C, const float* gamma, const float* beta,
                                const float* save_mean, const float* save_invstd, float slope, float* dx, float* dgamma, float* dbeta,
                                double* workspace, void* cuda_stream);
/* Training step, row f-1 (second slice): the gradient of a stride-1 convolution with respect to its INPUT (dgrad) on the tcgen05
 * implicit-GEMM kernel of the forward path.  Replaces what Keras / TensorFlow run through cuDNN for the backward of every stride-1
 * Conv2D of the reference's model when it trains (src/space/yolov3_detect.py:206-211 via src/space/face_detection.py:361-381,
 * :602-630): dX = conv(dY, flip(W)^T), i.e. a 'same' convolution of the output gradient with the weights transposed and both filter
 * axes reversed.  A single-convolution handle holds one k x k (k = 1 or 3), stride-1, zero-padded convolution cin -> cout over
 * [batch][height][width] maps without bias or activation: bf16 operands (the caller's float32 tensors are rounded once), float32
 * accumulation, float32 output.
 *   fvy_conv_create      : cin a multiple of 32 (the kernel's K chunk), cout <= 1024.  For dgrad of a forward layer Ci -> Co create
 *                          the handle with cin = Co, cout = Ci.  stride = 2 (forward only): a 3 x 3 filter with the reference's
 *                          ZeroPadding2D(1) + 'valid' geometry over an even-sized height x width INPUT (yolov3_detect.py:204-211),
 *                          cin a multiple of 64; y is then [batch][height/2][width/2][cout].
 *   fvy_conv_set_weights : w = the torch / Keras-transposed weight tensor [Co][Ci][k][k] float32 in DEVICE memory.  dgrad = 0: this
 *                          handle computes the forward convolution (cin = Ci, cout = Co); dgrad = 1: it computes dX from dY.
 *   fvy_conv_run         : x [batch][height][width][cin] float32 NHWC (DEVICE), y [batch][height][width][cout] float32 NHWC (DEVICE);
 *                          everything is enqueued on `cuda_stream` (the caller's stream), nothing synchronises.
 * Only fvy_conv_set_weights / fvy_conv_run / fvy_destroy / fvy_launch_count apply to such a handle. */
int fvy_conv_create(int device, int height, int width, int cin, int cout, int ksize, int stride, int max_batch, fvy_handle** out);
int fvy_conv_set_weights(fvy_handle* h, const float* w_dev, int dgrad, void* cuda_stream);
int fvy_conv_run(fvy_handle* h, const float* x_dev, int batch, float* y_dev, void* cuda_stream);
/* Training step, row f-1 (third slice): the gradient of a stride-1, zero-padded k x k convolution (k = 1 or 3) with respect to its
 * WEIGHTS - the other half of what cuDNN does for the reference's Conv2D layers when it trains (same lines as above):
 *   dW[co][ci][r][s] = sum over batch, y, x of dY[n][y][x][co] * X[n][y + r - k/2][x + s - k/2][ci]      (zero outside the map)
 * x [batch][height][width][cin] and dy [batch][height][width][cout] are float32 NHWC tensors in DEVICE memory (rounded once to
 * bf16; fp32 accumulation), dw [cout][cin][k][k] float32 (the torch / Keras-transposed layout) is overwritten.  cin and cout must be
 * multiples of 64.  x_scratch / dy_scratch: DEVICE buffers of fvy_conv_wgrad_scratch_rows(max batch, height, width) * cin (resp.
 * cout) * 2 bytes that the caller ZEROES ONCE and then only hands to this function for that (height, width, channels) - they hold
 * the bf16 operands in the shared-halo pixel-major form, whose halo pixels and margins stay zero.  dw_work: DEVICE scratch of the size
 * of dw (the tcgen05 kernel accumulates per tap, [k*k][cout][cin]; may be NULL for k = 1).  Enqueued on `cuda_stream`.
 * cout a multiple of 128: wgrad_tc_kernel (tcgen05.mma with MN-major operand descriptors); otherwise conv_wgrad_kernel (mma.sync).
 * stride = 2 (k = 3, cout a multiple of 128): height / width are dY's map, x is [batch][2 height][2 width][cin] and x_scratch holds four
 * phase planes: 4 * fvy_conv_wgrad_scratch_rows(batch, ...) * cin * 2 bytes, spaced by that (batch-dependent) row count - hand a stride-2
 * x_scratch to calls of ONE batch size only (zero it again before using it with another). */
long long fvy_conv_wgrad_scratch_rows(int batch, int height, int width);
int fvy_conv_wgrad(const float* x_dev, const float* dy_dev, int batch, int height, int width, int cin, int cout, int ksize, int stride,
                   void* x_scratch, void* dy_scratch, float* dw_dev, float* dw_work, void* cuda_stream);
/* Pre-processing of FaceDetector.evaluate / FaceDetector.test (src/space/face_detection.py:657-690 and :798-835):
 *   image = imread(file) / 255;  image = cv.resize(image, (w_p, h_p), interpolation=cv.INTER_CUBIC);
 *   image = cv.copyMakeBorder(image, pad_t, pad_b, pad_l, pad_r, cv.BORDER_CONSTANT, value=[0, 0, 0])
 * for one uint8 [src_h][src_w][3] image (host or device memory), on the GPU: OpenCV's separable bicubic (a = -0.75, float
 * coefficients, float64 accumulation, horizontal pass first, replicated borders, source coordinate (d + 0.5) * scale - 0.5)
 * followed by the cast to float32 that Keras applies to its input.  The result - bit-identical to float32(reference image) -
 * is written to image slot `index` (0 <= index < max_batch) of the handle's staged batch, a DEVICE buffer
 * float32 [max_batch][net_h][net_w][3] returned by fvy_staged_images(); pass that pointer as `images` (dtype FVY_F32) to
 * fvy_detect / fvy_forward.  w_p, h_p, pad_t, pad_l are the caller's (the reference computes them with Python float
 * arithmetic, face_detection.py:664-688); rows / columns outside [pad_t, pad_t + h_p) x [pad_l, pad_l + w_p) are zero. */
int fvy_letterbox_u8(fvy_handle* h, const unsigned char* src, int src_h, int src_w, int w_p, int h_p, int pad_t, int pad_l, int index);
float* fvy_staged_images(fvy_handle* h);
/* Copies the first `batch` staged images to dst (host or device float32 [batch][net_h][net_w][3]); synchronous. */
int fvy_read_staged(fvy_handle* h, int batch, float* dst);
/* Pinned host memory helpers (cudaHostAlloc / cudaFreeHost). */
void* fvy_host_alloc(size_t bytes);
void fvy_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* FVY_H_ */
