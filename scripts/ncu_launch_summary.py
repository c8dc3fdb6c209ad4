"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list.

    python scripts/ncu_launch_summary.py gpurun_out/launches.csv [--last-step N_LAUNCHES] \
        [--csv profiles/rX_ncu_launch_list.csv] [--json profiles/ncu_traffic_by_class.json]

Per kernel (template instantiation) and per bench.py kernel class: launches, summed duration, share of the step and the
DRAM bytes actually moved per launch (dram read + write) - the `roofline.traffic` figure bench.py reports next to the
algorithmic bytes.  ncu times are cold-cache and serialised: only the SHARES are comparable with the CUDA-event numbers.
"""
import argparse
import collections
import csv
import json
import re


def bench_class(name):
    """ncu kernel name -> the class (C-ABI entry point / GEMM kind) that bench.py aggregates by"""
    m = re.match(r"(?:void )?gemm_tc_kernel<\(?(?:int\))?(\d+), \(?(?:\w+\))?(\d+), \(?(?:\w+\))?(\d+), \(?(?:\w+\))?(\d+)[,>]", name)
    if m:
        # classes by operand layout, like ops.gemm_raw: A MN-major = weight gradient (dY^T X), B MN-major = data gradient
        # (dY W, the untouched [N,K] weight; also the few attention-style products with a transposed B), else forward
        a_mn, b_mn = int(m.group(2)), int(m.group(3))
        return "gemm_tc_wgrad" if a_mn else ("gemm_tc_dgrad" if b_mn else "gemm_tc_fwd")
    table = [("attn_kernel<0>", "cmx_attn_fwd"), ("attn_kernel<1>", "cmx_attn_bwd"), ("attn_kernel<(int)0>", "cmx_attn_fwd"),
             ("attn_kernel<(int)1>", "cmx_attn_bwd"), ("ln_bwd", "cmx_layernorm_bwd"), ("ln_fwd", "cmx_layernorm_fwd"),
             ("adamw_flat", "cmx_adamw_flat"), ("col2im_nhwc", "cmx_col2im_nhwc"), ("im2col_nhwc", "cmx_im2col_nhwc"),
             ("im2col_nchw", "cmx_im2col_nchw"), ("colsum", "cmx_colsum"), ("bn_bwd_apply", "cmx_bn_bwd_apply"),
             ("bn_bwd_reduce", "cmx_bn_bwd_reduce"), ("bn_apply", "cmx_bn_apply"), ("colstats", "cmx_colstats"),
             ("upsample_bwd", "cmx_upsample_bwd_multi"), ("upsample_sum", "cmx_upsample_sum_fwd"),
             ("ce_upsampled", "cmx_ce_upsampled_fwd_bwd"), ("frm_rectify_bwd", "cmx_frm_rectify_bwd"),
             ("frm_rectify_fwd", "cmx_frm_rectify_fwd"), ("pool_avgmax_bwd", "cmx_pool_avgmax_bwd"),
             ("pool_partial", "cmx_pool_avgmax_fwd"), ("pool_finalize", "cmx_pool_avgmax_fwd"),
             ("smallm_linear_fwd", "cmx_smallm_linear_fwd"), ("smallm_d", "cmx_smallm_linear_bwd"), ("smallm_bwd_fused", "cmx_smallm_linear_bwd"),
             ("cast_f32_bf16", "cmx_cast_f32_bf16"), ("relu_bwd", "cmx_relu_bwd"), ("softmax_dim2_fwd", "cmx_softmax_dim2_fwd"),
             ("softmax_dim2_bwd", "cmx_softmax_dim2_bwd")]
    m = re.match(r"(?:void )?dwconv_(?:tiled|tma)_kernel<\(?(?:int\))?(\d+), \(?(?:int\))?(\d+)(?:, \(?(?:int\))?\d+)?>", name)
    if m:
        return {0: "cmx_dwconv3x3_fwd", 1: "cmx_dwconv3x3_bwd_pre", 2: "cmx_dwconv3x3_dgrad"}[int(m.group(2))]
    for key, cls in table:
        if key in name:
            return cls
    return "torch:" + re.sub(r"<.*", "", name.replace("void ", ""))[:48] if "at::" in name else re.sub(r"\(.*", "", name)[:48]


def parse(path):
    """-> list of {name, us, rd, wr} in launch order"""
    rows = list(csv.reader(l for l in open(path, errors="replace") if not l.startswith("==")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    ki, ni, vi, ii, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID"), h.index("Metric Unit")
    launches = collections.OrderedDict()
    scale_t = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
    scale_b = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        d = launches.setdefault(r[ii], {"name": r[ki], "us": 0.0, "rd": None, "wr": None})
        if r[ni] == "gpu__time_duration.sum":
            d["us"] = v * scale_t.get(r[ui], 1e-3)
        elif r[ni] == "dram__bytes_read.sum":
            d["rd"] = v * scale_b.get(r[ui], 1.0)
        elif r[ni] == "dram__bytes_write.sum":
            d["wr"] = v * scale_b.get(r[ui], 1.0)
    return list(launches.values())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("path")
    ap.add_argument("--last-step", type=int, default=0, help="keep only the last N launches (one training step)")
    ap.add_argument("--from-last", default="", help="keep the launches from the last kernel whose name contains this marker "
                    "(convw_pack_multi = first-but-one launch of a training step)")
    ap.add_argument("--csv")
    ap.add_argument("--json")
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    ls = parse(a.path)
    if a.last_step:
        ls = ls[-a.last_step:]
    if a.from_last:
        idx = [i for i, l in enumerate(ls) if a.from_last in l["name"]]
        if idx:
            ls = ls[max(idx[-1] - 1, 0):]   # the flat weight cast precedes the conv-weight pack
    total = sum(l["us"] for l in ls)
    by_class = collections.OrderedDict()
    for l in ls:
        c = by_class.setdefault(bench_class(l["name"]), {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "have": 0})
        c["n"] += 1
        c["us"] += l["us"]
        if l["rd"] is not None:
            c["rd"] += l["rd"]
            c["wr"] += l["wr"] or 0.0
            c["have"] += 1
    lines = ["# %s" % a.note, "# %d launches, %.1f us summed kernel time (cold-cache, serialised: compare SHARES)" % (len(ls), total),
             "class,launches,total_us,share,avg_us,dram_read_MB_per_launch,dram_write_MB_per_launch,dram_GBps"]
    out = {}
    for k, c in sorted(by_class.items(), key=lambda kv: -kv[1]["us"]):
        have = c["have"] > 0
        per = (c["rd"] + c["wr"]) / c["n"] if have else None
        lines.append("%s,%d,%.1f,%.4f,%.2f,%s,%s,%s" % (
            k, c["n"], c["us"], c["us"] / total if total else 0, c["us"] / c["n"],
            "%.3f" % (c["rd"] / c["n"] / 1e6) if have else "", "%.3f" % (c["wr"] / c["n"] / 1e6) if have else "",
            "%.1f" % ((c["rd"] + c["wr"]) / (c["us"] * 1e-6) / 1e9) if have and c["us"] else ""))
        out[k] = {"launches": c["n"], "share": c["us"] / total if total else 0, "dram_bytes_per_launch": per}
    text = "\n".join(lines) + "\n"
    print(text)
    if a.csv:
        open(a.csv, "w").write(text)
    if a.json:
        json.dump({"source": a.note, "by_class": out}, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
