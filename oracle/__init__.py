"""ORACLE — test infrastructure only.

CPU restatements of the reference hot path used as the parity checker:
  * ``postproc_c.c`` / ``postproc.py`` — decode / letterbox / IoU / NMS (plain C + ctypes wrapper)
  * ``darknet_ref.py``                — torch-CPU fp32 restatement of the Keras graph
  * ``ref_loader.py``                 — stub-imports the REAL reference post-processing from
                                        /root/reference (build container only; used to pin the
                                        restatement and to generate tests/golden/)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this package.  The product (``face_vijnana_yolov3_b200``) never does.
"""
