"""Turns the ncu CSV logs of tools/evidence_v6.sh into the small files kept under profiles/.

  launches  <ncu --metrics gpu__time_duration.sum csv> <out.csv>   launch list + one-step share by kernel
  traffic   <ncu dram bytes csv> <out.csv> <out.json>               per-launch DRAM bytes of the LAST forward in the log
  full      <ncu --page raw csv> <out.csv>                          key metrics of a --set full capture, one line per launch
"""
import csv, json, sys, collections


def read_ncu(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    return rows[hi], rows[hi + 1:]


def launches(src, dst):
    h, rows = read_ncu(src)
    iN, iV, iG, iB = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
    L = [(r[iN].split("(")[0].replace("void ", "").replace("fvy::", ""), r[iG], r[iB], float(r[iV].replace(",", "")) / 1e3) for r in rows if len(r) > iV]
    stems = [i for i, l in enumerate(L) if l[0].startswith("stem_")]
    a, b = stems[1], stems[2]            # second step of the run (the first is the warm-up)
    share = collections.OrderedDict()
    for l in L[a:b]:
        share.setdefault(l[0], [0, 0.0]); share[l[0]][0] += 1; share[l[0]][1] += l[3]
    tot = sum(v[1] for v in share.values())
    conv = sum(v[1] for k, v in share.items() if k.startswith(("conv_", "stem_")))
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv  python bench.py --steps 2 --warmup 1   (cold-cache, serialised: compare shares)\n")
        f.write(f"# one step (launches {a}..{b - 1}): {b - a} launches, {tot:.1f} us; conv stack (stem + conv kernels) {100 * conv / tot:.1f} %, decode + NMS {100 * (tot - conv) / tot:.1f} %\n")
        f.write("# share by kernel: " + "; ".join(f"{k} x{v[0]} {v[1]:.1f} us ({100 * v[1] / tot:.1f}%)" for k, v in share.items()) + "\n")
        f.write("id,kernel,grid,block,us\n")
        for i, l in enumerate(L):
            f.write(f'{i},"{l[0]}","{l[1]}","{l[2]}",{l[3]:.3f}\n')
    print(open(dst).read().split("\n")[1])


def traffic(src, dst_csv, dst_json):
    h, rows = read_ncu(src)
    iI, iN, iM, iV = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    per = collections.OrderedDict()
    for r in rows:
        if len(r) <= iV: continue
        d = per.setdefault(int(r[iI]), {"k": r[iN].split("(")[0].replace("void ", "").replace("fvy::", "")})
        d[r[iM]] = float(r[iV].replace(",", ""))
    ids = list(per)
    stems = [i for i in ids if per[i]["k"].startswith("stem_")]
    sel = [i for i in ids if i >= stems[-1]]          # the last forward
    unit = {r[iM]: r[h.index("Metric Unit")] for r in rows if len(r) > iV}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = wr = t = 0.0
    with open(dst_csv, "w") as f:
        f.write("id,kernel,dram_read_bytes,dram_write_bytes,time_ns\n")
        for n, i in enumerate(sel):
            d = per[i]
            r_ = d["dram__bytes_read.sum"] * scale[unit["dram__bytes_read.sum"]]; w_ = d["dram__bytes_write.sum"] * scale[unit["dram__bytes_write.sum"]]
            tn = d["gpu__time_duration.sum"] * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}[unit["gpu__time_duration.sum"]]
            rd += r_; wr += w_; t += tn
            f.write(f'{n},"{d["k"]}",{r_:.0f},{w_:.0f},{tn:.0f}\n')
    import hashlib, os
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "face_vijnana_yolov3_b200", "libfvy.so")
    sha = hashlib.sha256(open(so, "rb").read()).hexdigest()[:16] if os.path.exists(so) else None
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hs = hashlib.sha256()          # the library's sources: nvcc output is not byte-reproducible, bench.py matches builds by this hash
    for name in sorted(os.listdir(os.path.join(root, "face_vijnana_yolov3_b200", "csrc"))):
        if name.endswith((".cu", ".cuh", ".h", ".inl")):
            hs.update(name.encode()); hs.update(open(os.path.join(root, "face_vijnana_yolov3_b200", "csrc", name), "rb").read())
    hs.update(open(os.path.join(root, "include", "fvy.h"), "rb").read())
    src_sha = hs.hexdigest()[:16]
    js = {"source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none, one forward ({len(sel)} launches: stem_strip_kernel, "
                    "conv_igemm_kernel and conv_chain_kernel), batch 40 @416, tools/evidence_r02.sh",
          "libfvy_sha16": sha, "libfvy_src_sha16": src_sha, "batch": 40, "net": 416,
          "launches": len(sel), "dram_bytes_read_per_step": rd, "dram_bytes_write_per_step": wr, "dram_bytes_per_step": rd + wr,
          "dram_bytes_per_launch": (rd + wr) / len(sel), "algorithmic_bytes_per_step_unfused": 7640332544, "ncu_time_sum_us": t / 1e3,
          "note": "writes that are still resident in the 126 MB L2 when a kernel ends are counted as reads of the next kernel or not at all; the sum over the forward is the meaningful figure"}
    json.dump(js, open(dst_json, "w"), indent=1)
    print(json.dumps(js)[:400])


KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__cluster_size", "smsp__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def full(src, dst):
    rows = list(csv.reader(open(src, errors="replace")))
    h, units, body = rows[0], rows[1], rows[2:]
    iN = h.index("Kernel Name")
    cols = [(i, c) for i, c in enumerate(h) if any(c == k or c.startswith(k) for k in KEYS) or "tensor" in c]
    with open(dst, "w") as f:
        f.write("# key metrics of an `ncu --set full --clock-control none --import-source on` capture (tools/evidence_v6.sh); one line per launch\n")
        f.write("kernel," + ",".join(f"{c} [{units[i]}]" for i, c in cols) + "\n")
        for r in body:
            f.write('"' + r[iN].split("(")[0].replace("void ", "") + '",' + ",".join(r[i] for i, _ in cols) + "\n")
    print(len(body), "launches,", len(cols), "metrics")


if __name__ == "__main__":
    {"launches": launches, "traffic": traffic, "full": full}[sys.argv[1]](*sys.argv[2:])
