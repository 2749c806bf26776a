"""Layer table of the Darknet-53 / YOLOv3 conv stack (host-side, pure Python).

Follows the graph the reference declares in
``src/space/yolov3_detect.py:196-311`` (``_conv_block`` / ``make_yolov3_model``) and the
truncated backbone + 3x3x6 head of ``src/space/face_detection.py:341-352,384-600``.

The table is *data*, not a graph builder: every conv is one record saying where its input
comes from, which tensor (if any) is added after the activation, and where its output goes.
The CUDA plan builder (``csrc/fvy_plan.cu``) holds the same table in C; ``tests/`` cross-check
the two through ``fvy_layer_info``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

HEAD_YOLO3 = 0  # three 1x1 heads, C = 3*(5+nb_class)  (yolov3_detect.py:277-308)
HEAD_FD6 = 1    # Darknet-53 base + one 3x3 'same' conv with bb_info_c_size outputs (face_detection.py:348-352)

YOLO_HEAD_IDX = (81, 93, 105)


@dataclass
class ConvSpec:
    idx: int            # Darknet cfg layer index (conv_<idx>, bnorm_<idx>)
    cin: int
    cout: int
    k: int              # 1 or 3
    stride: int         # 1 or 2
    bn: bool            # BatchNormalization(eps=1e-3) + no conv bias; else conv bias
    leaky: bool         # LeakyReLU(0.1)
    src: int            # idx of the conv (or add) whose output feeds this conv; -1 = network input;
                        # -2 = concat A (up(conv_84), skip_61); -3 = concat B (up(conv_96), skip_36)
    res: Optional[int]  # idx of the conv whose *post-add* output is added after the activation
    level: int          # log2 of the spatial down-sampling of the OUTPUT (0 => H, 5 => H/32)
    name: str = field(default="")

    @property
    def n_kernel(self) -> int:
        return self.cout * self.cin * self.k * self.k

    @property
    def n_params(self) -> int:
        return self.n_kernel + (4 * self.cout if self.bn else self.cout)


def _block(specs: List[ConvSpec], convs, src: int, level: int, skip: bool) -> int:
    """Append one ``_conv_block`` (yolov3_detect.py:196-215). Returns idx of the block output."""
    x = src
    skip_src = None
    for count, (idx, cin, cout, k, s, bn, leaky) in enumerate(convs):
        if skip and count == len(convs) - 2:
            skip_src = x                       # tensor entering the second-to-last conv (:201-202)
        if s == 2:
            level += 1
        last = count == len(convs) - 1
        specs.append(ConvSpec(idx, cin, cout, k, s, bn, leaky, x,
                              skip_src if (skip and last) else None, level, f"conv_{idx}"))
        x = idx
    return x


def yolo3_table(nb_class: int = 1) -> List[ConvSpec]:
    """All 75 convs of ``make_yolov3_model`` with C = 3*(5+nb_class) head channels."""
    C = 3 * (5 + nb_class)
    T, F = True, False
    s: List[ConvSpec] = []
    x = _block(s, [(0, 3, 32, 3, 1, T, T), (1, 32, 64, 3, 2, T, T), (2, 64, 32, 1, 1, T, T), (3, 32, 64, 3, 1, T, T)], -1, 0, True)
    x = _block(s, [(5, 64, 128, 3, 2, T, T), (6, 128, 64, 1, 1, T, T), (7, 64, 128, 3, 1, T, T)], x, 1, True)
    x = _block(s, [(9, 128, 64, 1, 1, T, T), (10, 64, 128, 3, 1, T, T)], x, 2, True)
    x = _block(s, [(12, 128, 256, 3, 2, T, T), (13, 256, 128, 1, 1, T, T), (14, 128, 256, 3, 1, T, T)], x, 2, True)
    for i in range(7):
        x = _block(s, [(16 + 3 * i, 256, 128, 1, 1, T, T), (17 + 3 * i, 128, 256, 3, 1, T, T)], x, 3, True)
    # skip_36 = x (:245)
    x = _block(s, [(37, 256, 512, 3, 2, T, T), (38, 512, 256, 1, 1, T, T), (39, 256, 512, 3, 1, T, T)], x, 3, True)
    for i in range(7):
        x = _block(s, [(41 + 3 * i, 512, 256, 1, 1, T, T), (42 + 3 * i, 256, 512, 3, 1, T, T)], x, 4, True)
    # skip_61 = x (:257)
    x = _block(s, [(62, 512, 1024, 3, 2, T, T), (63, 1024, 512, 1, 1, T, T), (64, 512, 1024, 3, 1, T, T)], x, 4, True)
    for i in range(3):
        x = _block(s, [(66 + 3 * i, 1024, 512, 1, 1, T, T), (67 + 3 * i, 512, 1024, 3, 1, T, T)], x, 5, True)
    x = _block(s, [(75, 1024, 512, 1, 1, T, T), (76, 512, 1024, 3, 1, T, T), (77, 1024, 512, 1, 1, T, T),
                   (78, 512, 1024, 3, 1, T, T), (79, 1024, 512, 1, 1, T, T)], x, 5, False)
    _block(s, [(80, 512, 1024, 3, 1, T, T), (81, 1024, C, 1, 1, F, F)], x, 5, False)            # yolo_82
    _block(s, [(84, 512, 256, 1, 1, T, T)], x, 5, False)                                         # -> up -> cat A
    x = _block(s, [(87, 768, 256, 1, 1, T, T), (88, 256, 512, 3, 1, T, T), (89, 512, 256, 1, 1, T, T),
                   (90, 256, 512, 3, 1, T, T), (91, 512, 256, 1, 1, T, T)], -2, 4, False)
    _block(s, [(92, 256, 512, 3, 1, T, T), (93, 512, C, 1, 1, F, F)], x, 4, False)               # yolo_94
    _block(s, [(96, 256, 128, 1, 1, T, T)], x, 4, False)                                         # -> up -> cat B
    _block(s, [(99, 384, 128, 1, 1, T, T), (100, 128, 256, 3, 1, T, T), (101, 256, 128, 1, 1, T, T),
               (102, 128, 256, 3, 1, T, T), (103, 256, 128, 1, 1, T, T), (104, 128, 256, 3, 1, T, T),
               (105, 256, C, 1, 1, F, F)], -3, 3, False)                                         # yolo_106
    return s


FD6_HEAD_IDX = 1000  # synthetic index for FaceDetector's 'output' conv (face_detection.py:348-352)


def fd6_table(bb_info_c_size: int = 6) -> List[ConvSpec]:
    """Darknet-53 base conv_0..conv_73 (52 convs, 23 adds) + 3x3 linear head with bias."""
    s = [c for c in yolo3_table(1) if c.idx <= 73]
    s.append(ConvSpec(FD6_HEAD_IDX, 1024, bb_info_c_size, 3, 1, False, False, 73, None, 5, "output"))
    return s


def table(head: int, nb_class: int = 1, bb_info_c_size: int = 6) -> List[ConvSpec]:
    return yolo3_table(nb_class) if head == HEAD_YOLO3 else fd6_table(bb_info_c_size)


def n_params(specs: List[ConvSpec]) -> int:
    return sum(c.n_params for c in specs)


def macs(specs: List[ConvSpec], h: int, w: int) -> int:
    """Algorithmic MACs per image: sum Ho*Wo*Cout*k*k*Cin, un-padded channel counts."""
    return sum((h >> c.level) * (w >> c.level) * c.cout * c.k * c.k * c.cin for c in specs)
