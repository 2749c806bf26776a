"""Minimal HDF5 reader / writer for Keras 2.2.4 weight files.

The reference (de)serialises its models with Keras ``model.save`` / ``load_model`` (src/space/face_detection.py:329, :337, :394,
:598, :630; src/space/yolov3_detect.py:573, :585): ``face_detector.h5`` and ``yolov3_base.h5`` are HDF5 files whose
``model_weights`` group holds one group per layer and, below it, one contiguous float32 dataset per weight
(``conv_12/conv_12/kernel:0``, ``bnorm_12/bnorm_12/gamma:0`` ...; for the FaceDetector the Darknet-53 base is a nested model, so
its weights sit one level deeper: ``model_weights/model_1/conv_12/kernel:0``).  h5py is not available in this image, so this module
restates the part of the published HDF5 file format (HDF5 File Format Specification, version 0 superblock - what h5py writes
with its default ``libver='earliest'`` and what Keras 2.2.4-era files are) that such files use:

    superblock v0 -> root symbol-table entry -> object header v1 (+ continuation blocks) -> symbol-table message
    -> group B-tree v1 ("TREE", any depth) -> symbol-table nodes ("SNOD") + local heap ("HEAP") names
    -> dataset object headers: dataspace (v1 / v2), datatype (IEEE float / fixed-point, little endian), data layout v3
       (contiguous or compact).  Chunked / compressed datasets and new-style ("OHDR") object headers are rejected loudly.

The writer produces the same subset (one SNOD per group: the superblock's leaf K is raised so that every group fits one node)
and exists so that (a) trained weights can be handed back to a Keras user as ``face_detector.h5`` and (b) the reader has files to
be tested on offline (tests/test_h5lite.py also checks the byte layout against hand-derived offsets of the specification).
"""
from __future__ import annotations

import struct
from typing import Dict, List, Tuple

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(ValueError):
    pass


# ----------------------------------------------------------------------------------------------------------------------
# reader
# ----------------------------------------------------------------------------------------------------------------------
class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        if buf[:8] != SIGNATURE:
            # the superblock may sit at 512, 1024, ... (user block); Keras files have none
            raise H5Error("not an HDF5 file (signature missing at offset 0)")
        ver = buf[8]
        if ver not in (0, 1):
            raise H5Error(f"superblock version {ver} (new-style file, libver='latest') is not supported: re-save with h5py's default libver")
        self.size_off, self.size_len = buf[13], buf[14]
        if (self.size_off, self.size_len) != (8, 8):
            raise H5Error("only 8-byte offsets / lengths are supported")
        p = 16
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", buf, p); p += 4
        p += 4                                     # file consistency flags
        if ver == 1:
            p += 4                                 # indexed storage internal node K + reserved
        self.base, _free, self.eof, _drv = struct.unpack_from("<QQQQ", buf, p); p += 32
        # root group symbol table entry
        _name_off, self.root_header, cache_type = struct.unpack_from("<QQI", buf, p)
        self.root_scratch = struct.unpack_from("<QQ", buf, p + 24) if cache_type == 1 else None

    # -- object headers ------------------------------------------------------------------------------------------------
    def messages(self, addr: int) -> List[Tuple[int, bytes]]:
        b = self.b
        addr += self.base
        if b[addr:addr + 4] == b"OHDR":
            raise H5Error("version-2 object headers (libver='latest') are not supported")
        ver, _res, nmsg, _refs, hsize = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise H5Error(f"object header version {ver} at {addr}")
        out: List[Tuple[int, bytes]] = []
        blocks = [(addr + 16, hsize)]              # the 12-byte prefix is padded to 16
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, p)
                data = b[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x0010:                # continuation
                    coff, clen = struct.unpack_from("<QQ", data, 0)
                    blocks.append((coff + self.base, clen))
                out.append((mtype, data))
        return out

    # -- groups ----------------------------------------------------------------------------------------------------------
    def _heap_name(self, heap_addr: int, off: int) -> str:
        b = self.b
        heap_addr += self.base
        if b[heap_addr:heap_addr + 4] != b"HEAP":
            raise H5Error("local heap signature missing")
        data_addr, = struct.unpack_from("<Q", b, heap_addr + 24)
        s = data_addr + self.base + off
        e = b.index(b"\x00", s)
        return b[s:e].decode("utf-8")

    def _btree_entries(self, node: int, heap: int, out: List[Tuple[str, int]]):
        b = self.b
        a = node + self.base
        if b[a:a + 4] == b"SNOD":
            n, = struct.unpack_from("<H", b, a + 6)
            for i in range(n):
                e = a + 8 + 40 * i
                name_off, header = struct.unpack_from("<QQ", b, e)
                out.append((self._heap_name(heap, name_off), header))
            return
        if b[a:a + 4] != b"TREE":
            raise H5Error(f"group B-tree node signature missing at {a}")
        ntype, _level, used = struct.unpack_from("<BBH", b, a + 4)
        if ntype != 0:
            raise H5Error("not a group B-tree")
        p = a + 24                                  # after the two sibling pointers
        for i in range(used):
            child, = struct.unpack_from("<Q", b, p + 8 + 16 * i)   # key_i (8) child_i (8) ... key_used
            self._btree_entries(child, heap, out)

    def children(self, header: int) -> Dict[str, int] | None:
        """name -> object header address if the object is an (old-style) group, else None."""
        for mtype, data in self.messages(header):
            if mtype == 0x0011:
                btree, heap = struct.unpack_from("<QQ", data, 0)
                out: List[Tuple[str, int]] = []
                self._btree_entries(btree, heap, out)
                return dict(out)
            if mtype in (0x0002, 0x0006):
                raise H5Error("new-style groups (link messages) are not supported")
        return None

    # -- datasets --------------------------------------------------------------------------------------------------------
    def dataset(self, header: int) -> np.ndarray:
        shape = dtype = None
        layout = None
        for mtype, data in self.messages(header):
            if mtype == 0x0001:
                ver, rank, flags = data[0], data[1], data[2]
                p = 8 if ver == 1 else 4
                shape = struct.unpack_from("<" + "Q" * rank, data, p) if rank else ()
            elif mtype == 0x0003:
                cls, bits0 = data[0] & 0x0F, data[1]
                size, = struct.unpack_from("<I", data, 4)
                if bits0 & 1:
                    raise H5Error("big-endian datasets are not supported")
                if cls == 1 and size in (2, 4, 8):
                    dtype = np.dtype(f"<f{size}")
                elif cls == 0 and size in (1, 2, 4, 8):
                    dtype = np.dtype(("<i" if bits0 & 8 else "<u") + str(size))
                else:
                    raise H5Error(f"datatype class {cls} size {size} is not supported")
            elif mtype == 0x0008:
                ver, lclass = data[0], data[1]
                if ver != 3:
                    raise H5Error(f"data layout message version {ver} is not supported")
                if lclass == 1:
                    addr, nbytes = struct.unpack_from("<QQ", data, 2)
                    layout = ("contiguous", addr, nbytes)
                elif lclass == 0:
                    n, = struct.unpack_from("<H", data, 2)
                    layout = ("compact", data[4:4 + n], n)
                else:
                    raise H5Error("chunked (compressed?) datasets are not supported: Keras writes contiguous float32 weights")
            elif mtype == 0x000B:
                raise H5Error("filtered datasets are not supported")
        if shape is None or dtype is None or layout is None:
            raise H5Error("object is not a dataset")
        count = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if layout[0] == "compact":
            raw = layout[1]
        else:
            if layout[1] == UNDEF:
                return np.zeros(shape, dtype)      # never written
            raw = self.b[layout[1] + self.base:layout[1] + self.base + count * dtype.itemsize]
        if len(raw) < count * dtype.itemsize:
            raise H5Error("dataset extends past the end of the file")
        return np.frombuffer(raw, dtype, count).reshape(shape).copy()

    def walk(self, header: int, prefix: str, out: Dict[str, np.ndarray]):
        kids = self.children(header)
        if kids is None:
            out[prefix] = self.dataset(header)
            return
        for name, h in kids.items():
            self.walk(h, f"{prefix}/{name}" if prefix else name, out)


def read_datasets(path: str) -> Dict[str, np.ndarray]:
    """Every dataset of the file as {'group/.../name': array}."""
    with open(path, "rb") as f:
        r = _Reader(f.read())
    out: Dict[str, np.ndarray] = {}
    r.walk(r.root_header, "", out)
    return out


# ----------------------------------------------------------------------------------------------------------------------
# writer (same subset)
# ----------------------------------------------------------------------------------------------------------------------
def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


class _Writer:
    LEAF_K = 512          # up to 1024 symbols per SNOD: every group of a Keras weight file fits one node

    def __init__(self):
        self.buf = bytearray(b"\x00" * 96)        # superblock v0 (56 bytes + 40-byte root entry)

    def alloc(self, data: bytes) -> int:
        self.buf += b"\x00" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    def _header(self, msgs: List[Tuple[int, bytes]]) -> int:
        body = b"".join(struct.pack("<HHB3x", t, len(_pad8(d)), 0) + _pad8(d) for t, d in msgs)
        return self.alloc(struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body)

    def dataset(self, arr: np.ndarray) -> int:
        arr = np.ascontiguousarray(arr)
        if arr.dtype.kind == "f":
            size = arr.dtype.itemsize
            exp_bits, man_bits, bias = {2: (5, 10, 15), 4: (8, 23, 127), 8: (11, 52, 1023)}[size]
            # class 1 v1; bit field: little endian, pad 0, mantissa normalisation 2 (implied msb), sign at the top bit
            dt = struct.pack("<BBBBI", 0x11, 0x20, size * 8 - 1, 0, size) + struct.pack("<HHBBBBI", 0, size * 8, man_bits, exp_bits, 0, man_bits, bias)
            arr = arr.astype(f"<f{size}")
        elif arr.dtype.kind in "iu":
            size = arr.dtype.itemsize
            dt = struct.pack("<BBBBI", 0x10, 0x08 if arr.dtype.kind == "i" else 0, 0, 0, size) + struct.pack("<HH", 0, size * 8)
            arr = arr.astype(("<i" if arr.dtype.kind == "i" else "<u") + str(size))
        else:
            raise H5Error(f"dtype {arr.dtype} is not supported")
        raw = arr.tobytes()
        data_addr = self.alloc(raw) if raw else UNDEF
        space = struct.pack("<BBB5x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape)
        layout = struct.pack("<BBQQ", 3, 1, data_addr, len(raw))
        return self._header([(0x0001, space), (0x0003, dt), (0x0008, layout)])

    def group(self, kids: Dict[str, int]) -> Tuple[int, int, int]:
        """-> (object header address, B-tree address, heap address)"""
        names = sorted(kids)                       # symbol-table entries are ordered by name
        if len(names) > 2 * self.LEAF_K:
            raise H5Error("too many entries for a single symbol-table node")
        heap_data = bytearray(b"\x00" * 8)         # offset 0: the empty string
        offs = []
        for n in names:
            offs.append(len(heap_data))
            heap_data += _pad8(n.encode("utf-8") + b"\x00")
        free_off = len(heap_data)
        heap_data += struct.pack("<QQ", 1, 16)     # one free block: next = 1 (none), size 16
        data_addr = self.alloc(bytes(heap_data))
        heap = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, data_addr))
        snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
        for n, o in zip(names, offs):
            snod += struct.pack("<QQII16x", o, kids[n], 0, 0)
        snod += b"\x00" * (40 * (2 * self.LEAF_K - len(names)))
        snod_addr = self.alloc(snod)
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, UNDEF, UNDEF)
        tree += struct.pack("<QQQ", 0, snod_addr, offs[-1] if offs else 0)
        tree += b"\x00" * (16 * (2 * 16 - 1))       # room for the node's 2K (internal K = 16) entries
        tree_addr = self.alloc(tree)
        header = self._header([(0x0011, struct.pack("<QQ", tree_addr, heap))])
        return header, tree_addr, heap

    def build(self, tree: dict) -> Tuple[int, int, int]:
        kids = {}
        for name, v in tree.items():
            if "/" in name:
                raise H5Error("nested names must be nested dicts")
            kids[name] = self.build(v)[0] if isinstance(v, dict) else self.dataset(np.asarray(v))
        return self.group(kids)

    def finish(self, root: Tuple[int, int, int]) -> bytes:
        header, btree, heap = root
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.LEAF_K, 16, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, header, 1, 0) + struct.pack("<QQ", btree, heap)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write_h5(path: str, tree: dict) -> None:
    """tree: nested dicts (groups) whose leaves are arrays (datasets)."""
    w = _Writer()
    root = w.build(tree)
    with open(path, "wb") as f:
        f.write(w.finish(root))


# ----------------------------------------------------------------------------------------------------------------------
# Keras weight layout <-> Darknet stream order
# ----------------------------------------------------------------------------------------------------------------------
def _layer_weights(datasets: Dict[str, np.ndarray]) -> Dict[str, Dict[str, np.ndarray]]:
    """{'conv_12': {'kernel': a, 'bias': b}, 'bnorm_12': {'gamma': ..}} from dataset paths ending in '<layer>/<weight>:0'
    (optimizer_weights are ignored)."""
    out: Dict[str, Dict[str, np.ndarray]] = {}
    for path, arr in datasets.items():
        parts = path.split("/")
        if len(parts) < 2 or parts[0] == "optimizer_weights":
            continue
        layer, w = parts[-2], parts[-1].split(":")[0]
        out.setdefault(layer, {})[w] = arr
    return out


def keras_h5_to_stream(path: str, specs) -> np.ndarray:
    """Weights of a Keras 2.2.4 .h5 (``model.save`` or ``save_weights``) in the Darknet stream order ``fvy_load_weights`` takes
    (yolov3_detect.py:91-119: per conv beta, gamma, mean, var then the kernel as (Cout, Cin, kh, kw); heads: bias then kernel).
    The FaceDetector's 3x3x6 head is the layer named 'output' (face_detection.py:348-352)."""
    from . import arch
    lw = _layer_weights(read_datasets(path))
    chunks = []
    for s in specs:
        conv = "output" if s.idx == arch.FD6_HEAD_IDX else f"conv_{s.idx}"
        if conv not in lw or "kernel" not in lw[conv]:
            raise H5Error(f"{path}: no kernel for layer {conv}")
        k = np.asarray(lw[conv]["kernel"], np.float32)
        if k.shape != (s.k, s.k, s.cin, s.cout):
            raise H5Error(f"{path}: {conv}/kernel has shape {k.shape}, expected {(s.k, s.k, s.cin, s.cout)}")
        if s.bn:
            bn = lw.get(f"bnorm_{s.idx}")
            if bn is None:
                raise H5Error(f"{path}: no bnorm_{s.idx}")
            for name in ("beta", "gamma", "moving_mean", "moving_variance"):
                v = np.asarray(bn[name], np.float32)
                if v.shape != (s.cout,):
                    raise H5Error(f"{path}: bnorm_{s.idx}/{name} has shape {v.shape}")
                chunks.append(v)
        else:
            chunks.append(np.asarray(lw[conv]["bias"], np.float32).reshape(s.cout))
        chunks.append(np.ascontiguousarray(k.transpose(3, 2, 0, 1)).reshape(-1))       # (kh, kw, Cin, Cout) -> (Cout, Cin, kh, kw)
    return np.concatenate(chunks).astype(np.float32)


def stream_to_keras_h5(path: str, stream: np.ndarray, specs, nested_base: str | None = None) -> None:
    """The inverse: a ``model_weights`` tree in Keras 2.2.4's layout.  ``nested_base``: name of the nested Darknet-53 base model
    (FaceDetector: its conv / bnorm weights live under model_weights/<nested_base>/, the head under model_weights/output/output/)."""
    from . import arch
    stream = np.asarray(stream, np.float32)
    mw: dict = {}
    base: dict = {}
    off = 0
    for s in specs:
        layers = {}
        if s.bn:
            beta, gamma, mean, var = (stream[off + i * s.cout:off + (i + 1) * s.cout] for i in range(4))
            off += 4 * s.cout
            layers[f"bnorm_{s.idx}"] = {"gamma:0": gamma, "beta:0": beta, "moving_mean:0": mean, "moving_variance:0": var}
            bias = None
        else:
            bias = stream[off:off + s.cout]; off += s.cout
        n = s.cout * s.cin * s.k * s.k
        kern = stream[off:off + n].reshape(s.cout, s.cin, s.k, s.k).transpose(2, 3, 1, 0); off += n
        conv = "output" if s.idx == arch.FD6_HEAD_IDX else f"conv_{s.idx}"
        layers[conv] = {"kernel:0": np.ascontiguousarray(kern)}
        if bias is not None:
            layers[conv]["bias:0"] = bias
        for name, w in layers.items():
            if nested_base is not None and name != "output":
                base[name] = w                                   # model_weights/<base>/<layer>/<weight>
            else:
                mw[name] = {name: w}                             # model_weights/<layer>/<layer>/<weight>
    if off != stream.size:
        raise H5Error("stream length does not match the layer table")
    if nested_base is not None:
        mw[nested_base] = base
    write_h5(path, {"model_weights": mw})
