"""Drop-in for the reference's ``src/space/face_detection.py`` FaceDetector, backed by libfvy.so.

    FaceDetector(conf)                      reference :312-382   (conf = face_vijnana_yolov3.json['fd_conf'])
      .detect(image) -> [BoundBox]          reference :885-949   (predict + decode + do_nms_v2 + select, one GPU call)
      .detect_batch(images) -> [[BoundBox]] additive: the same per image, batched
      .evaluate() / .test()                 reference :632-781 / :783-883  (host image loop kept; per-image detect on the GPU)
      .train()                              reference :602-630   (data-parallel step of train.py: NCCL gradient all-reduce)
    main()                                  reference :951-985

The model is the reference's: Darknet-53 base conv_0..conv_73 (:384-600) + Conv2D(6, 3x3, same, linear)
(:348-352) -> (B,13,13,6) ``[obj, bx, by, bw, bh, cls]``.  Weights: the reference's own Keras files - ``face_detector.h5``
(``model_loading``, :329/:337) and ``yolov3_base.h5`` (``yolov3_base_model_load``, :394) - are read with ``h5lite`` (a minimal
HDF5 reader: h5py is not needed), a Darknet ``yolov3.weights`` file fills the backbone (:398-402), a flat
``face_detector.fvyw`` stream (``FaceDetector.save_weights``) is the native format, else Keras-default random initialisation
as in the reference's untrained model.  ``train()`` writes both ``face_detector.fvyw`` and a Keras-layout ``face_detector.h5``.
"""
from __future__ import annotations

import glob
import json
import os
import platform
import time

import numpy as np

from .. import _lib as L
from .. import arch, synth
from ..engine import Engine, post_params
from .yolov3_detect import BoundBox, WeightReader

DEBUG = True


class FaceDetector(object):
    """Face detector to use yolov3 (same constants as the reference, :66-73)."""

    MODEL_PATH = 'face_detector.h5'
    WEIGHTS_PATH = 'face_detector.fvyw'
    CELL_SIZE = 13

    def __init__(self, conf, device=0, max_batch=None):
        self.conf = conf
        self.raw_data_path = self.conf.get('raw_data_path')
        self.hps = self.conf['hps']
        self.nn_arch = self.conf['nn_arch']
        self.model_loading = self.conf.get('model_loading', False)
        self.cell_image_size = self.nn_arch['image_size'] // self.CELL_SIZE      # :325
        if self.nn_arch['image_size'] % 32:
            raise ValueError("nn_arch.image_size must be a multiple of 32")
        self.device = device
        self.max_batch = int(max_batch or 1)
        self.specs = arch.fd6_table(self.nn_arch['bb_info_c_size'])
        self._stream = self._initial_stream()
        self._engine = None
        self._sharded = None

    # ------------------------------------------------------------------ weights
    def _initial_stream(self) -> np.ndarray:
        n = arch.n_params(self.specs)
        if self.model_loading:
            if os.path.exists(self.WEIGHTS_PATH):
                s = np.fromfile(self.WEIGHTS_PATH, dtype='<f4')
                if s.size != n:
                    raise ValueError(f"{self.WEIGHTS_PATH} holds {s.size} floats, model needs {n}")
                return s
            if os.path.exists(self.MODEL_PATH):                                  # load_model(self.MODEL_PATH), :329 / :337
                from .. import h5lite
                return h5lite.keras_h5_to_stream(self.MODEL_PATH, self.specs)
            raise FileNotFoundError(f"model_loading is set but neither {self.WEIGHTS_PATH} nor {self.MODEL_PATH} exists")
        stream = synth.darknet_stream(self.specs, 0, synth.INIT_KERAS_DEFAULT)   # Keras default init of the untrained model
        if self.conf.get('yolov3_base_model_load') and os.path.exists('yolov3_base.h5'):   # load_model('yolov3_base.h5'), :393-396
            from .. import h5lite
            base = [c for c in self.specs if c.idx <= 73]
            stream[:arch.n_params(base)] = h5lite.keras_h5_to_stream('yolov3_base.h5', base)
        elif os.path.exists('yolov3.weights'):                                   # YOLOV3Base, :398-402
            wr = WeightReader('yolov3.weights')
            n_base = arch.n_params([c for c in self.specs if c.idx <= 73])
            stream[:n_base] = wr.read_bytes(n_base)
        return stream

    def save_weights(self, path=None):
        np.asarray(self._stream, '<f4').tofile(path or self.WEIGHTS_PATH)

    def save_model(self, path=None):
        """``self.model.save(self.MODEL_PATH)`` (:630): the weights as a Keras 2.2.4 ``model_weights`` tree (nested Darknet-53 base
        'model_1' + 'output' head), readable by Keras' ``load_weights`` and by this class (``model_loading``)."""
        from .. import h5lite
        h5lite.stream_to_keras_h5(path or self.MODEL_PATH, self._stream, self.specs, nested_base='model_1')

    def set_weight_stream(self, stream):
        stream = np.ascontiguousarray(stream, np.float32)
        if stream.size != arch.n_params(self.specs):
            raise ValueError("wrong weight stream length")
        self._stream = stream
        if self._engine is not None:
            self._engine.load_weights(stream)
        if self._sharded is not None:
            self._sharded.load_weights(stream)

    @property
    def engine(self) -> Engine:
        if self._engine is None:
            s = self.nn_arch['image_size']
            self._engine = Engine(s, s, head=L.HEAD_FD6, max_batch=self.max_batch, device=self.device,
                                  bb_info_c_size=self.nn_arch['bb_info_c_size'])
            self._engine.load_weights(self._stream)
        return self._engine

    def _pp(self):
        return post_params(obj_thresh=self.hps['face_conf_th'], nms_thresh=self.hps['nms_iou_th'],
                           num_cands=self.hps['num_cands'], arith=L.ARITH_F64)

    # ------------------------------------------------------------------ detect (:885-949)
    @staticmethod
    def _to_boxes(dets, n):
        out = []
        for d in dets[:n]:
            sc = np.float32(d['score'])
            out.append(BoundBox(np.int64(d['xmin']), np.int64(d['ymin']), np.int64(d['xmax']), np.int64(d['ymax']),
                                objness=np.float32(d['objness']), classes=[sc]))
        return out

    def detect(self, image):
        """image: (1, S, S, 3) float in [0,1] -> face candidate BoundBoxes, ascending score, <= num_cands."""
        image = np.ascontiguousarray(image)
        if image.dtype not in (np.float32, np.float64):
            image = image.astype(np.float64)
        if image.shape[0] != 1:
            raise ValueError("detect takes one image (1,S,S,3); use detect_batch for more")
        dets, counts = self.engine.detect(image, pp=self._pp())
        return self._to_boxes(dets[0], int(counts[0]))

    def detect_batch(self, images):
        """The same per image for a whole batch.  With ``conf['multi_gpu']`` and ``conf['num_gpus'] > 1`` the batch is sharded over
        that many devices in this process (``shard.ShardedDetector``: the one-line replacement of the reference's
        ``multi_gpu_model(self.model, gpus=num_gpus)``, :330, :369) - identical results, image order kept."""
        images = np.ascontiguousarray(images)
        if images.dtype not in (np.float32, np.float64, np.uint8):
            images = images.astype(np.float64)
        n_gpus = int(self.conf.get('num_gpus', 1)) if self.conf.get('multi_gpu') else 1
        if n_gpus > 1:
            from ..shard import ShardedDetector
            per = -(-images.shape[0] // n_gpus)
            if self._sharded is None or self._sharded.engines[0].max_batch < per:
                if self._sharded is not None:
                    self._sharded.close()
                s = self.nn_arch['image_size']
                self._sharded = ShardedDetector(list(range(n_gpus)), s, s, head=L.HEAD_FD6, max_batch_per_device=per,
                                                bb_info_c_size=self.nn_arch['bb_info_c_size'])
                self._sharded.load_weights(self._stream)
            dets, counts = self._sharded.detect(images, pp=self._pp())
            return [self._to_boxes(dets[b], int(counts[b])) for b in range(images.shape[0])]
        if images.shape[0] > self.max_batch:
            self.max_batch = int(images.shape[0])
            if self._engine is not None:
                self._engine.close()
                self._engine = None
        dets, counts = self.engine.detect(images, pp=self._pp())
        return [self._to_boxes(dets[b], int(counts[b])) for b in range(images.shape[0])]

    # ------------------------------------------------------------------ host image loop (:645-735, :788-883)
    def _letterbox_geom(self, w, h):
        """Resized extent and top / left padding exactly as the reference computes them (:664-688): Python float arithmetic."""
        S = self.nn_arch['image_size']
        pad_t = pad_l = 0
        if w >= h:
            w_p, h_p = S, int(h / w * S)
            pad_t = (S - h_p) // 2
        else:
            h_p, w_p = S, int(w / h * S)
            pad_l = (S - w_p) // 2
        return w_p, h_p, pad_t, pad_l

    def _letterbox(self, image):
        import cv2 as cv
        S = self.nn_arch['image_size']
        w, h = image.shape[1], image.shape[0]
        pad_t = pad_b = pad_l = pad_r = 0
        if w >= h:
            w_p, h_p = S, int(h / w * S)
            pad = S - h_p
            pad_t, pad_b = pad // 2, pad // 2 + (pad % 2)
            image = cv.resize(image, (w_p, h_p), interpolation=cv.INTER_CUBIC)
            image = cv.copyMakeBorder(image, pad_t, pad_b, 0, 0, cv.BORDER_CONSTANT, value=[0, 0, 0])
        else:
            h_p, w_p = S, int(w / h * S)
            pad = S - w_p
            pad_l, pad_r = pad // 2, pad // 2 + (pad % 2)
            image = cv.resize(image, (w_p, h_p), interpolation=cv.INTER_CUBIC)
            image = cv.copyMakeBorder(image, 0, 0, pad_l, pad_r, cv.BORDER_CONSTANT, value=[0, 0, 0])
        return image[np.newaxis, :], (w, h, pad_t, pad_l)

    def _unletterbox(self, boxes, geom):
        S = self.nn_arch['image_size']
        w, h, pad_t, pad_l = geom
        for box in boxes:
            if w >= h:
                box.xmin = np.min([box.xmin * w / S, w])
                box.xmax = np.min([box.xmax * w / S, w])
                box.ymin = np.min([np.max([box.ymin - pad_t, 0]) * w / S, h])
                box.ymax = np.min([np.max([box.ymax - pad_t, 0]) * w / S, h])
            else:
                box.xmin = np.min([np.max([box.xmin - pad_l, 0]) * h / S, w])
                box.xmax = np.min([np.max([box.xmax - pad_l, 0]) * h / S, w])
                box.ymin = np.min([box.ymin * h / S, h])
                box.ymax = np.min([box.ymax * h / S, h])

    def _run_files(self, draw_dir=None):
        """The reference's per-file loop (:645-735, :788-883) with the pre-processing and the detector call on the GPU: each
        file's pixels go to the device as uint8 and are letterboxed there (``fvy_letterbox_u8``: image/255, cv.resize INTER_CUBIC,
        zero border - bit-identical to the reference's host code, which `_letterbox` keeps for comparison), ``max_batch`` of them
        go through ONE detect call, and rows are written in the reference's file order.  ``max_batch = 1`` is the reference."""
        import cv2 as cv
        test_path = self.conf['test_path']
        output_file_path = self.conf['output_file_path']
        file_names = glob.glob(os.path.join(test_path, '*.jpg'))
        bs = max(1, int(self.max_batch))
        eng = self.engine
        with open(output_file_path, 'w') as f:
            for start in range(0, len(file_names), bs):
                chunk = file_names[start:start + bs]
                originals, geoms = [], []
                for i, file_name in enumerate(chunk):
                    if DEBUG:
                        print(start + i + 1, '/', len(file_names), file_name)
                    image_o = np.ascontiguousarray(cv.imread(file_name, cv.IMREAD_COLOR)[:, :, ::-1])   # RGB like skimage.io.imread
                    h, w = image_o.shape[0], image_o.shape[1]
                    w_p, h_p, pad_t, pad_l = self._letterbox_geom(w, h)
                    eng.letterbox(image_o, i, w_p, h_p, pad_t, pad_l)
                    originals.append(image_o); geoms.append((w, h, pad_t, pad_l))
                dets, counts = eng.detect(eng.staged(len(chunk)), pp=self._pp())
                all_boxes = [self._to_boxes(dets[b], int(counts[b])) for b in range(len(chunk))]
                for file_name, image_o, geom, boxes in zip(chunk, originals, geoms, all_boxes):
                    self._unletterbox(boxes, geom)
                    base = file_name.split('\\')[-1] if platform.system() == 'Windows' else file_name.split('/')[-1]
                    for count, box in enumerate(boxes, 1):
                        if count > 60:                                                # :729, :870
                            break
                        f.write(base + ',' + str(box.xmin) + ',' + str(box.ymin) + ',')
                        f.write(str(box.xmax - box.xmin) + ',' + str(box.ymax - box.ymin) + ',' + str(box.get_score()) + '\n')
                    if draw_dir is not None and len(boxes) > 0:
                        canvas = np.ascontiguousarray(image_o[:, :, ::-1])
                        for box in boxes:
                            cv.rectangle(canvas, (int(box.xmin), int(box.ymin)), (int(box.xmax), int(box.ymax)), (0, 255, 0), 2)
                        cv.imwrite(os.path.join(draw_dir, base[:-4] + '_detected.jpg'), canvas)

    def evaluate(self):
        """Detect on every ``test_path/*.jpg``, write ``output_file_path`` CSV and ``results/*_detected.jpg``."""
        import shutil
        res = os.path.join(self.conf['test_path'], 'results')
        if os.path.isdir(res):
            shutil.rmtree(res)
        os.mkdir(res)
        self._run_files(draw_dir=res)

    def test(self):
        """Detect on every ``test_path/*.jpg`` and write the ``file,x,y,w,h,score`` CSV (<= 60 rows per file)."""
        self._run_files(draw_dir=None)

    def train(self, device=None, on_step=None):
        """Train on ``raw_data_path/training.csv`` (reference :602-630): MSE against the 13x13x6 ground truth, Keras Adam,
        ``hps['epochs']`` x ``hps['step']`` steps.  Under ``torchrun`` (one process per GPU, ``conf['multi_gpu']``) every
        rank takes its contiguous slice of each batch and gradients are all-reduced over NCCL - the replacement of
        ``multi_gpu_model`` (:330, :369).  Weights are saved to ``WEIGHTS_PATH`` and loaded into the inference engine."""
        import torch
        import torch.distributed as dist
        from .. import train as T
        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
        if device is None:
            device = f"cuda:{self.device}" if torch.cuda.is_available() else "cpu"
        seq = T.TrainingSequence(self.raw_data_path, self.hps, self.nn_arch, self.CELL_SIZE)
        # hps['fvy_conv_mode'] (optional, not a reference key; default 0 = the reference's fp32 arithmetic): bits 1 / 2 / 4 put the
        # stride-1 convolutions' dgrad / wgrad / forward on this repo's bf16 tensor-core kernels (train._ConvFn)
        trainer = T.DataParallelTrainer(self.hps, device=device, bb_info_c_size=self.nn_arch['bb_info_c_size'], stream=self._stream,
                                        fvy_conv_mode=int(self.hps.get('fvy_conv_mode', 0)))
        for epoch in range(self.hps['epochs']):
            for i in range(len(seq)):
                images, gts = seq[i]
                xs, ts = T.slice_for_rank(images, gts, rank, world)      # a short last batch may leave a rank without images: it
                loss = trainer.step(torch.from_numpy(np.ascontiguousarray(xs)), torch.from_numpy(np.ascontiguousarray(ts)),
                                    global_batch=len(images))             # contributes zeros and still joins the all-reduce
                if on_step is not None:
                    on_step(epoch, i, loss)
                elif DEBUG and rank == 0:
                    print(f"epoch {epoch + 1}/{self.hps['epochs']} step {i + 1}/{len(seq)} loss {loss:.6f}")
        stream = trainer.weight_stream()
        self.set_weight_stream(stream)
        if rank == 0:
            print('Save the model.')
            self.save_weights()
            self.save_model()


def main():
    """Same dispatch as the reference (:951-985): reads face_vijnana_yolov3.json['fd_conf']."""
    name = "face_vijnana_yolov3_win.json" if platform.system() == 'Windows' else "face_vijnana_yolov3.json"
    with open(name, 'r') as f:
        conf = json.load(f)['fd_conf']
    fd = FaceDetector(conf)
    ts = time.time()
    if conf['mode'] == 'train':
        fd.train()
    elif conf['mode'] == 'evaluate':
        fd.evaluate()
    elif conf['mode'] == 'test':
        fd.test()
    te = time.time()
    print('Elasped time: {0:f}s'.format(te - ts))


if __name__ == '__main__':
    main()
