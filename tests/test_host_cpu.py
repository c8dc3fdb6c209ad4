"""CPU tests of the host side: C-ABI library exports, module tree / state_dict contract of the EncoderDecoder
drop-in, loud failure without a GPU, pickling, and the N>1 gradient hand-off under DDP (gloo)."""
import ctypes
import os
import pickle
import subprocess
import sys

import pytest
import torch
import torch.nn as nn

from oracle import cmx_ref
from rgbx_semantic_segmentation_b200 import _lib
from rgbx_semantic_segmentation_b200.models.builder import EncoderDecoder

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Cfg:
    backbone = "mit_b2"
    decoder = "MLPDecoder"
    decoder_embed_dim = 512
    num_classes = 9
    pretrained_model = None
    bn_eps = 1e-3
    bn_momentum = 0.1


class CfgB0(Cfg):
    backbone = "mit_b0"
    decoder_embed_dim = 64


def test_library_exports_every_declared_symbol():
    if not _lib.lib_available():
        import __graft_entry__
        __graft_entry__.build()
    protos = _lib.parse_header()
    assert len(protos) >= 40
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(lib, name), name
    lib.cmx_version.restype = ctypes.c_int
    assert lib.cmx_version() == 100                      # no compute call without a GPU
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and "cmx_" in l}
    assert exported == set(protos), exported ^ set(protos)


def test_state_dict_schema_and_param_groups():
    m = EncoderDecoder(Cfg, nn.CrossEntropyLoss(reduction="mean", ignore_index=255), nn.BatchNorm2d)
    sd = m.state_dict()
    sch = cmx_ref.state_dict_schema(cmx_ref.MIT_SPECS["mit_b2"], 9)
    assert list(sd.keys()) == list(sch.keys()) and len(sd) == 837
    assert all(tuple(sd[k].shape) == sch[k][0] for k in sch)
    assert sum(p.numel() for p in m.parameters()) == 66565521
    # utils/init_func.py:33-57 group_weight: 288 decay + 522 no-decay tensors
    decay, no_decay = [], []
    for mod in m.modules():
        if isinstance(mod, (nn.Linear, nn.Conv2d)):
            decay.append(mod.weight)
            if mod.bias is not None:
                no_decay.append(mod.bias)
        elif isinstance(mod, (nn.BatchNorm2d, nn.LayerNorm)):
            no_decay += [mod.weight, mod.bias]
    assert (len(decay), len(no_decay)) == (288, 522)
    # train-built => decoder BN eps/momentum overridden (App. A-4); FFM BNs stay at 1e-5
    assert m.decode_head.linear_fuse[1].eps == 1e-3 and m.backbone.FFMs[0].channel_emb.norm.eps == 1e-5
    e = EncoderDecoder(Cfg, None, nn.BatchNorm2d)
    assert e.decode_head.linear_fuse[1].eps == 1e-5
    assert m.aux_head is None and m.channels == [64, 128, 320, 512] and m.cfg is Cfg
    # drop-path schedule incl. the stage-2 quirk (App. A-6)
    rgb_p, ext_p = cmx_ref.drop_path_probs(cmx_ref.MIT_SPECS["mit_b2"])
    for s in range(4):
        for i, blk in enumerate(getattr(m.backbone, f"block{s + 1}")):
            assert abs(getattr(blk.drop_path, "drop_prob", 0.0) - rgb_p[s][i]) < 1e-12
        for i, blk in enumerate(getattr(m.backbone, f"extra_block{s + 1}")):
            assert abs(getattr(blk.drop_path, "drop_prob", 0.0) - ext_p[s][i]) < 1e-12


def test_b4_channels_fixed_and_unsupported_configs_raise():
    c = type("C", (Cfg,), {"backbone": "mit_b4", "num_classes": 5})
    m = EncoderDecoder(c, None, nn.BatchNorm2d)
    assert m.decode_head.linear_c4.proj.in_features == 512        # reference bug App. A-1 not reproduced
    assert sum(p.numel() for p in m.parameters()) == 139856269
    for bad in ({"backbone": "swin_s"}, {"decoder": "UPernet"}, {"feature_rectify_module": "IFRM"}):
        with pytest.raises(NotImplementedError):
            EncoderDecoder(type("C", (Cfg,), bad), None, nn.BatchNorm2d)


def test_cpu_inputs_fail_loudly_and_model_pickles():
    c = CfgB0
    m = EncoderDecoder(c, nn.CrossEntropyLoss(ignore_index=255), nn.BatchNorm2d)
    z = torch.zeros(1, 3, 32, 32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(z, z)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(z, z, torch.zeros(1, 32, 32, dtype=torch.long))
    with pytest.raises(NotImplementedError):
        EncoderDecoder(c, nn.MSELoss(), nn.BatchNorm2d)(z, z, torch.zeros(1, 32, 32, dtype=torch.long))
    m2 = pickle.loads(pickle.dumps(m))                              # engine/evaluator.py:131 pickles the model
    assert list(m2.state_dict().keys()) == list(m.state_dict().keys())
    # after the first step the parameters are views of ONE flat buffer; a pickle must not carry that buffer per tensor
    m._eng()._flatten(torch.device("cpu"))
    n_bytes = sum(p.numel() * 4 for p in m.parameters())
    assert next(m.parameters()).untyped_storage().nbytes() >= n_bytes
    blob = pickle.dumps(m)
    assert len(blob) < 2 * n_bytes + (4 << 20), len(blob)
    m3 = pickle.loads(blob)
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m3.state_dict().values()))
    assert m3._engine is None and next(m3.parameters()).untyped_storage().nbytes() < (1 << 20)
    assert next(m.parameters()).untyped_storage().nbytes() >= n_bytes          # the live model keeps its flat views


def test_graph_cache_is_least_recently_used_and_bounded(monkeypatch):
    monkeypatch.setenv("CMX_MAX_GRAPHS", "3")
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    m = EncoderDecoder(CfgB0, None, nn.BatchNorm2d)
    for k in "abc":
        assert m._graph_entry(k) is None                 # first sight: warm-up entry, caller runs eagerly
    assert m._graph_entry("a") == {"warm": 1}            # hit -> most recently used
    assert m._graph_entry("d") is None and list(m._graphs) == ["c", "a", "d"]    # "b" was the least recently used
    assert m._graph_entry("b") is None and list(m._graphs) == ["a", "d", "b"]
    m._eng()._flatten(torch.device("cpu"))               # re-flattening drops every capture (raw pointers into the old buffers)
    assert m._graphs == {}


DDP_WORKER = r'''
import os, sys, torch, torch.nn as nn, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from rgbx_semantic_segmentation_b200.models.builder import EncoderDecoder
from rgbx_semantic_segmentation_b200.engine import Engine
class Cfg:
    backbone = "mit_b0"; decoder = "MLPDecoder"; decoder_embed_dim = 64; num_classes = 5
    pretrained_model = None; bn_eps = 1e-3; bn_momentum = 0.1
rank = int(os.environ["RANK"])
dist.init_process_group("gloo")
torch.manual_seed(0)
m = EncoderDecoder(Cfg, nn.CrossEntropyLoss(ignore_index=255), nn.BatchNorm2d)
eng = m._eng()
eng._flatten(torch.device("cpu"))
def fake_step(rgb, x, label):            # stands in for the CUDA step: rank-dependent gradient, rank-dependent loss
    eng.flat_g.fill_(float(rank + 1))
    return torch.tensor(float(rank))
m._run_step = fake_step
mode = sys.argv[2]
if mode == "ddp":
    ddp = torch.nn.parallel.DistributedDataParallel(m)
else:
    from rgbx_semantic_segmentation_b200.parallel import FlatDataParallel
    with torch.no_grad():
        for p in m.parameters():
            p.add_(float(rank))                      # desynchronise: the wrapper must broadcast rank 0's values
    ddp = FlatDataParallel(m)
    ref = [p.detach().clone() for p in m.parameters()]
    for p in ref:
        dist.broadcast(p, src=0)
    assert all(torch.equal(a, b) for a, b in zip(ref, m.parameters())), "parameters not broadcast from rank 0"
z = torch.zeros(1, 3, 32, 32)
loss = ddp(z, z, torch.zeros(1, 32, 32, dtype=torch.long))
loss.backward()
ok = all(torch.allclose(p.grad, torch.full_like(p.grad, 1.5)) for p in m.parameters())   # mean of 1 and 2
lt = loss.detach().clone(); dist.all_reduce(lt); lt /= dist.get_world_size()             # utils/pyt_utils.py:119-124
print("RESULT", rank, ok, float(lt), len(list(m.parameters())))
dist.destroy_process_group()
'''


@pytest.mark.parametrize("mode,port", [("ddp", "29533"), ("flat", "29534")])
def test_ddp_gradient_handoff_world2_gloo(tmp_path, mode, port):
    script = tmp_path / "w.py"
    script.write_text(DDP_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=port, WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, mode], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, o in enumerate(outs):
        line = [l for l in o.splitlines() if l.startswith("RESULT")]
        assert line, o[-2000:]
        _, rk, ok, lt, n = line[0].split()
        assert ok == "True" and abs(float(lt) - 0.5) < 1e-6, line


def test_sass_holds_tcgen05_and_tma_and_no_legacy_mma_outside_the_fallback():
    """static proof (no GPU): the GEMM / attention kernels in the built library issue tcgen05.mma (UTC*MMA), read their
    accumulators from tensor memory (LDTM) and move tiles with TMA (UTMALDG / UTMASTG); legacy mma.sync (HMMA) exists only
    in the generic strided fallback GEMM.  Same parser as scripts/sass_summary.py -> profiles/r1_sass_resource_summary.txt."""
    import csv
    import shutil
    import subprocess
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "scripts", "sass_summary.py")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    rows = list(csv.DictReader(l for l in out.stdout.splitlines() if not l.startswith("#")))
    by = {}
    for r in rows:
        by.setdefault(r["kernel"].split("<")[0], []).append(r)
    for fam in ("gemm_tc_kernel", "attn_kernel"):
        assert by[fam], fam
        for r in by[fam]:
            assert int(r["UTCMMA"]) > 0 and int(r["LDTM"]) > 0 and int(r["UTMALDG"]) > 0 and int(r["HMMA"]) == 0, r
    assert [k for k, v in by.items() if any(int(r["HMMA"]) for r in v)] == ["gemm_wmma_kernel"]


def test_dice_ce_mirror_equals_reference_golden(golden_dir):
    """utils.loss_opr.DiceCELoss (the torch formula the GPU tests use as checker for whole-model runs) == the reference class
    on the committed fixture (tests/golden/make_golden_dice.py)"""
    import numpy as np
    import torch
    from rgbx_semantic_segmentation_b200.utils.loss_opr import DiceCELoss
    z = np.load(os.path.join(golden_dir, "dice.npz"))
    for n in ("d5", "d9", "d40"):
        lg = torch.from_numpy(z[n + "_logits"]).requires_grad_(True)
        loss = DiceCELoss(alpha=float(z[n + "_meta"][1]))(lg, torch.from_numpy(z[n + "_target"]))
        g, = torch.autograd.grad(loss, lg)
        assert abs(loss.item() - float(z[n + "_loss"])) < 1e-6
        assert torch.allclose(g, torch.from_numpy(z[n + "_grad"]), rtol=1e-5, atol=1e-8)
