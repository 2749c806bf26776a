"""h5lite: the minimal HDF5 reader / writer for Keras 2.2.4 weight files.  h5py is not available offline, so the
byte layout is checked against offsets derived by hand from the HDF5 File Format Specification (version 0 superblock, version 1
object headers, symbol-table groups), and the reader is additionally exercised on a HAND-ASSEMBLED file the writer did not produce
(multi-node group B-tree, continuation block, compact dataset, v2 dataspace) - the shapes real h5py files take."""
import struct

import numpy as np
import pytest

from face_vijnana_yolov3_b200 import arch, h5lite, synth


def test_roundtrip_and_byte_layout(tmp_path):
    p = str(tmp_path / "a.h5")
    k = np.arange(2 * 3 * 4 * 5, dtype=np.float32).reshape(3, 4, 5, 2) / 7
    tree = {"model_weights": {"conv_0": {"conv_0": {"kernel:0": k}}, "bnorm_0": {"bnorm_0": {"gamma:0": np.ones(5, np.float32)}}},
            "ints": np.array([[1, -2], [3, 4]], np.int32)}
    h5lite.write_h5(p, tree)
    d = h5lite.read_datasets(p)
    assert set(d) == {"model_weights/conv_0/conv_0/kernel:0", "model_weights/bnorm_0/bnorm_0/gamma:0", "ints"}
    assert np.array_equal(d["model_weights/conv_0/conv_0/kernel:0"], k) and d["ints"].dtype == np.int32 and np.array_equal(d["ints"], tree["ints"])
    b = open(p, "rb").read()
    # superblock v0 (spec III.A): signature, versions, 8-byte offsets / lengths, EOF address, root symbol-table entry with cached B-tree / heap
    assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8] == 0 and b[13] == 8 and b[14] == 8
    assert struct.unpack_from("<Q", b, 40)[0] == len(b)
    root_hdr, cache = struct.unpack_from("<QI", b, 64)
    btree, heap = struct.unpack_from("<QQ", b, 80)
    assert cache == 1 and b[btree:btree + 4] == b"TREE" and b[heap:heap + 4] == b"HEAP"
    # root object header v1 (spec IV.A.1.a): version 1, one message = symbol table (0x0011) pointing at the same B-tree / heap
    ver, _, nmsg, refs, size = struct.unpack_from("<BBHII", b, root_hdr)
    mtype, msize = struct.unpack_from("<HH", b, root_hdr + 16)
    assert (ver, nmsg, refs, mtype, msize) == (1, 1, 1, 0x0011, 16) and struct.unpack_from("<QQ", b, root_hdr + 24) == (btree, heap)
    # group B-tree node -> one SNOD whose entries are sorted by name; names live in the local heap's data segment
    snod = struct.unpack_from("<Q", b, btree + 32)[0]
    assert b[snod:snod + 4] == b"SNOD" and struct.unpack_from("<H", b, snod + 6)[0] == 2
    seg = struct.unpack_from("<Q", b, heap + 24)[0]
    names = [b[seg + struct.unpack_from("<Q", b, snod + 8 + 40 * i)[0]:].split(b"\x00")[0] for i in range(2)]
    assert names == [b"ints", b"model_weights"]


def _hand_assembled_file():
    """A file laid out the way h5py (libver earliest, K = 4) writes them, assembled byte by byte here: root group with a TWO-level
    B-tree (internal node -> two leaf nodes -> SNODs), a dataset whose header continues in a second block, a compact dataset and a
    version-2 dataspace."""
    buf = bytearray(96)

    def put(data):
        while len(buf) % 8:
            buf.append(0)
        a = len(buf); buf.extend(data); return a

    def hdr(msgs, split=False):
        def enc(t, d):
            d = d + b"\x00" * (-len(d) % 8)
            return struct.pack("<HHB3x", t, len(d), 0) + d
        if not split:
            body = b"".join(enc(t, d) for t, d in msgs)
            return put(struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body)
        tail = b"".join(enc(t, d) for t, d in msgs[1:])
        ca = put(tail)
        body = enc(*msgs[0]) + enc(0x0010, struct.pack("<QQ", ca, len(tail)))
        return put(struct.pack("<BBHII4x", 1, 0, len(msgs) + 1, 1, len(body)) + body)

    f32 = struct.pack("<BBBBI", 0x11, 0x20, 31, 0, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    a = np.arange(6, dtype="<f4").reshape(2, 3) + 0.5
    da = put(a.tobytes())
    ds_a = hdr([(0x0001, struct.pack("<BBB5x", 1, 2, 0) + struct.pack("<QQ", 2, 3)), (0x0003, f32), (0x0008, struct.pack("<BBQQ", 3, 1, da, 24))], split=True)
    c = np.array([9.0, 8.0], "<f4")
    ds_c = hdr([(0x0001, struct.pack("<BBBB", 2, 1, 0, 1) + struct.pack("<Q", 2)), (0x0003, f32), (0x0008, struct.pack("<BBH", 3, 0, 8) + c.tobytes())])
    ds_z = hdr([(0x0001, struct.pack("<BBB5x", 1, 1, 0) + struct.pack("<Q", 4)), (0x0003, f32), (0x0008, struct.pack("<BBQQ", 3, 1, 0xFFFFFFFFFFFFFFFF, 16))])
    names = [b"alpha:0", b"beta:0", b"zeta:0"]
    heap_data = bytearray(8); offs = []
    for n in names:
        offs.append(len(heap_data)); heap_data += n + b"\x00" * (8 - len(n) % 8)
    seg = put(bytes(heap_data))
    heap = put(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), 0xFFFFFFFFFFFFFFFF, seg))

    def snod(entries):
        s = b"SNOD" + struct.pack("<BBH", 1, 0, len(entries))
        for o, h in entries:
            s += struct.pack("<QQII16x", o, h, 0, 0)
        return put(s + b"\x00" * (40 * (8 - len(entries))))
    s0 = snod([(offs[0], ds_a), (offs[1], ds_c)]); s1 = snod([(offs[2], ds_z)])

    def tree(level, kids, keys):
        t = b"TREE" + struct.pack("<BBHQQ", 0, level, len(kids), 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF)
        for i, k in enumerate(kids):
            t += struct.pack("<QQ", keys[i], k)
        t += struct.pack("<Q", keys[len(kids)])
        return put(t + b"\x00" * 512)
    l0 = tree(0, [s0], [0, offs[1]]); l1 = tree(0, [s1], [offs[1], offs[2]])
    top = tree(1, [l0, l1], [0, offs[1], offs[2]])
    root = hdr([(0x0011, struct.pack("<QQ", top, heap))])
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, 0xFFFFFFFFFFFFFFFF, len(buf), 0xFFFFFFFFFFFFFFFF) + struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", top, heap)
    buf[:96] = sb
    return bytes(buf), a, c


def test_reader_on_hand_assembled_file(tmp_path):
    data, a, c = _hand_assembled_file()
    p = tmp_path / "hand.h5"
    p.write_bytes(data)
    d = h5lite.read_datasets(str(p))
    assert set(d) == {"alpha:0", "beta:0", "zeta:0"}
    assert np.array_equal(d["alpha:0"], a) and np.array_equal(d["beta:0"], c) and np.array_equal(d["zeta:0"], np.zeros(4, np.float32))


def test_rejects_what_it_does_not_support(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not hdf5" * 20)
    with pytest.raises(h5lite.H5Error):
        h5lite.read_datasets(str(p))
    data, _, _ = _hand_assembled_file()
    bad = bytearray(data); bad[8] = 2                        # version-2 superblock (libver latest)
    p.write_bytes(bytes(bad))
    with pytest.raises(h5lite.H5Error):
        h5lite.read_datasets(str(p))


@pytest.mark.parametrize("nested", [None, "model_1"])
def test_keras_layout_to_darknet_stream_roundtrip(tmp_path, nested):
    """stream -> Keras-layout .h5 (kernels (kh,kw,Cin,Cout), BN gamma/beta/moving_*; FaceDetector: nested base + 'output' head)
    -> stream: identical; the kernel transpose and the beta,gamma,mean,var order are those of yolov3_detect.py:96-119."""
    specs = [s for s in arch.fd6_table(6) if s.idx <= 3 or s.idx == arch.FD6_HEAD_IDX]
    specs[-1].cin = 64                                            # a small head on top of conv_3 keeps the file tiny
    stream = synth.darknet_stream(specs, 5, synth.INIT_BN_EXERCISING)
    p = str(tmp_path / "fd.h5")
    h5lite.stream_to_keras_h5(p, stream, specs, nested_base=nested)
    d = h5lite.read_datasets(p)
    key = ("model_weights/model_1/conv_1/kernel:0" if nested else "model_weights/conv_1/conv_1/kernel:0")
    assert d[key].shape == (3, 3, 32, 64) and "model_weights/output/output/bias:0" in d
    # Darknet order of conv_1: beta, gamma, mean, var, kernel (Cout, Cin, kh, kw)
    off = specs[0].n_params
    beta = stream[off:off + 64]
    bkey = ("model_weights/model_1/bnorm_1/beta:0" if nested else "model_weights/bnorm_1/bnorm_1/beta:0")
    assert np.array_equal(d[bkey], beta)
    kern = stream[off + 256:off + 256 + 64 * 32 * 9].reshape(64, 32, 3, 3)
    assert np.array_equal(d[key], kern.transpose(2, 3, 1, 0))
    back = h5lite.keras_h5_to_stream(p, specs)
    assert np.array_equal(back, stream)


def test_facedetector_model_loading_reads_keras_h5(tmp_path, monkeypatch):
    """conf['model_loading'] with the reference's face_detector.h5 (face_detection.py:327-337) no longer raises: the file is read
    with h5lite and yields the stream it was written from (no GPU needed: the engine is created lazily)."""
    from face_vijnana_yolov3_b200.space.face_detection import FaceDetector
    monkeypatch.chdir(tmp_path)
    conf = {"mode": "test", "raw_data_path": "", "test_path": "", "output_file_path": "", "multi_gpu": False, "num_gpus": 1,
            "yolov3_base_model_load": False, "hps": {"face_conf_th": 0.5, "nms_iou_th": 0.5, "num_cands": 60},
            "nn_arch": {"image_size": 416, "bb_info_c_size": 6}, "model_loading": True}
    with pytest.raises(FileNotFoundError):
        FaceDetector(conf)
    specs = arch.fd6_table(6)
    stream = synth.darknet_stream(specs, 2, synth.INIT_BN_EXERCISING)
    h5lite.stream_to_keras_h5("face_detector.h5", stream, specs, nested_base="model_1")
    fd = FaceDetector(conf)
    assert np.array_equal(fd._stream, stream)
    fd.save_model("again.h5")
    assert np.array_equal(h5lite.keras_h5_to_stream("again.h5", specs), stream)
