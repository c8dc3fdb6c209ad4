"""CPU tests of the measurement plumbing: the `bench.py --impl reference` JSON line (the driver parses it), the ncu
launch-list summariser that feeds `roofline.traffic`, and the committed traffic file bench.py reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def test_reference_arm_line_matches_the_contract():
    """one bounded step of the reference's algorithm on the host cores; same metric / unit / workload as our arm"""
    env = dict(os.environ, OMP_NUM_THREADS=str(min(8, os.cpu_count() or 1)))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    import bench
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT
    # the reference arm is quoted on OUR arm's config (the dict both arms print), its bounded sample is described separately
    assert d["config"] == bench.workload_config(1) and d["config"]["workload"] == bench.WORKLOAD and d["sample"]
    assert set(d["config"]) == {"workload", "global_batch", "per_gpu_batch", "parallelism", "l2"}
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1
    assert d["value"] > 0 and abs(d["value"] - 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]
    cb = d["cpu_baseline"]
    # "reference" = the unmodified reference model from baseline/_ref (or /root/reference); "port" only when neither exists
    from baseline import ref_loader
    assert cb["kind"] == ("reference" if ref_loader.available() else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_ncu_kernel_names_map_to_bench_classes():
    import ncu_launch_summary as n
    cases = {
        "void gemm_tc_kernel<128, 1, 1, 0>(CUtensorMap_st, CUtensorMap_st)": "gemm_tc_wgrad",
        "void gemm_tc_kernel<(int)64, (int)0, (int)1, (int)0>(CUtensorMap_st)": "gemm_tc_dgrad",
        "void gemm_tc_kernel<64, 0, 1, 1, 4>(CUtensorMap_st)": "gemm_tc_dgrad",    # grouped data gradient (5 template arguments)
        "void gemm_tc_kernel<(int)128, (bool)0, (bool)0, (bool)1, (int)4>(CUtensorMap_st)": "gemm_tc_fwd",
        "void gemm_tc_kernel<128, 0, 0, 0>(CUtensorMap_st)": "gemm_tc_fwd",
        "void attn_kernel<1>(CUtensorMap_st, AttnArgs)": "cmx_attn_bwd",
        "void dwconv_tiled_kernel<2, 1>(const __nv_bfloat16 *, long)": "cmx_dwconv3x3_bwd_pre",
        "void dwconv_tiled_kernel<0, 2>(const __nv_bfloat16 *, long)": "cmx_dwconv3x3_dgrad",
        "void ln_bwd_v2_kernel<__nv_bfloat16, float, float, 32, 2>(const T1 *)": "cmx_layernorm_bwd",
        "void bn_bwd_apply_v8_kernel<__nv_bfloat16, float, float, __nv_bfloat16>(const T1 *)": "cmx_bn_bwd_apply",
        "upsample_bwd_multi_kernel(const __nv_bfloat16 *, int, int, int, UpDst, UpDst, UpDst, int)": "cmx_upsample_bwd_multi",
        "void at::native::vectorized_elementwise_kernel<4, at::native::CUDAFunctor_add<float>>(int)": "torch:at::native::vectorized_elementwise_kernel",
    }
    for name, cls in cases.items():
        assert n.bench_class(name) == cls, (name, n.bench_class(name))


def test_ncu_summary_parses_a_launch_list(tmp_path):
    import ncu_launch_summary as n
    p = tmp_path / "l.csv"
    hdr = '"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"\n'
    def row(i, name, metric, unit, val):
        return '"%d","1","python","h","%s","1","7","(256, 1, 1)","(10, 1, 1)","0","10.0","Command line profiler metrics","%s","%s","%s"\n' % (i, name, metric, unit, val)
    body = "==PROF== Connected\n" + hdr
    for i, (name, us, rd, wr) in enumerate([("convw_pack_multi_kernel(int)", "3,000", "1.5", "0.5"),
                                            ("void gemm_tc_kernel<128, 1, 1, 0>(X)", "20,000", "30", "1"),
                                            ("void gemm_tc_kernel<128, 1, 1, 0>(X)", "10,000", "10", "1")]):
        body += row(i, name, "gpu__time_duration.sum", "ns", us)
        body += row(i, name, "dram__bytes_read.sum", "Mbyte", rd)
        body += row(i, name, "dram__bytes_write.sum", "Mbyte", wr)
    p.write_text(body)
    ls = n.parse(str(p))
    assert [round(l["us"], 3) for l in ls] == [3.0, 20.0, 10.0]
    assert ls[1]["rd"] == 30e6 and ls[2]["wr"] == 1e6
    out = tmp_path / "t.json"
    sys.argv = ["ncu_launch_summary.py", str(p), "--json", str(out)]
    n.main()
    t = json.loads(out.read_text())["by_class"]
    assert t["gemm_tc_wgrad"]["launches"] == 2 and abs(t["gemm_tc_wgrad"]["dram_bytes_per_launch"] - 21e6) < 1


def test_committed_traffic_file_covers_the_dominant_classes():
    with open(os.path.join(ROOT, "profiles", "ncu_traffic_by_class.json")) as f:
        t = json.load(f)["by_class"]
    for k in ("gemm_tc_wgrad", "gemm_tc_fwd", "gemm_tc_dgrad", "cmx_layernorm_bwd", "cmx_dwconv3x3_bwd_pre", "cmx_attn_fwd"):
        assert t[k]["dram_bytes_per_launch"] > 0 and t[k]["launches"] > 0


def test_bench_has_no_undefined_names():
    """bench.py's GPU arm cannot run here; at least every name it loads must be bound somewhere (a typo in the line assembly
    would only surface on the GPU box)"""
    import ast
    import builtins
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    bound = set(dir(builtins)) | {"__file__", "__name__"}
    for node in ast.walk(tree):
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)):
            bound.add(node.name)
            if isinstance(node, ast.FunctionDef):
                a = node.args
                bound.update(x.arg for x in a.args + a.kwonlyargs + a.posonlyargs)
                bound.update(x.arg for x in (a.vararg, a.kwarg) if x)
        elif isinstance(node, ast.Lambda):
            bound.update(x.arg for x in node.args.args)
        elif isinstance(node, ast.Name) and isinstance(node.ctx, (ast.Store, ast.Del)):
            bound.add(node.id)
        elif isinstance(node, (ast.Import, ast.ImportFrom)):
            bound.update((al.asname or al.name).split(".")[0] for al in node.names)
        elif isinstance(node, ast.ExceptHandler) and node.name:
            bound.add(node.name)
        elif isinstance(node, ast.Global):
            bound.update(node.names)
    loads = {n.id for n in ast.walk(tree) if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load)}
    assert not (loads - bound), sorted(loads - bound)
