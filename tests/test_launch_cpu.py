"""The one-node launcher: every rank must receive exactly what the reference's Engine reads
(engine/engine.py:40-58,60-73): `--local_rank r` and `-p port` on argv, WORLD_SIZE / RANK / MASTER_ADDR in the environment."""
import json
import os
import sys
import time

from rgbx_semantic_segmentation_b200 import launch

# a stand-in for train.py whose parser spells its flags like the reference Engine's (engine.py:60-73) and which
# rendezvouses the way Engine.__init__ does (MASTER_PORT overwritten from --port, init_method env://) - on gloo here
_FAKE_TRAIN = r'''
import argparse, json, os, sys
import torch, torch.distributed as dist
p = argparse.ArgumentParser()
p.add_argument('-d', '--devices', default='')
p.add_argument('-c', '--continue', dest='continue_fpath')
p.add_argument('--local_rank', default=0, type=int)
p.add_argument('-p', '--port', type=str, default='16005', dest='port')
p.add_argument('--out')
a = p.parse_args()
world = int(os.environ['WORLD_SIZE'])
os.environ['MASTER_PORT'] = a.port
dist.init_process_group(backend='gloo', world_size=world, init_method='env://')
t = torch.tensor([float(a.local_rank + 1)])
dist.all_reduce(t)
json.dump(dict(local_rank=a.local_rank, rank=dist.get_rank(), world=world, sum=t.item(), devices=a.devices,
               addr=os.environ['MASTER_ADDR']), open(a.out + '.%d' % a.local_rank, 'w'))
dist.destroy_process_group()
'''


def test_rank_env_and_argv():
    env = launch.rank_env(3, 8, 29511, base={"PATH": "/bin", "WORLD_SIZE": "1"})
    assert env["RANK"] == env["LOCAL_RANK"] == "3" and env["WORLD_SIZE"] == "8"
    assert env["MASTER_ADDR"] == "127.0.0.1" and env["MASTER_PORT"] == "29511" and env["PATH"] == "/bin"
    argv = launch.rank_argv("train.py", ["-d", "0-7"], 3, 29511, python="py")
    assert argv == ["py", "-u", "train.py", "--local_rank", "3", "-p", "29511", "-d", "0-7"]


def test_launch_world2_gloo(tmp_path):
    script = tmp_path / "fake_train.py"
    script.write_text(_FAKE_TRAIN)
    out = str(tmp_path / "res")
    rc = launch.launch(str(script), ["-d", "0-1", "--out", out], nproc=2, port=29533)
    assert rc == 0
    res = [json.load(open(out + ".%d" % r)) for r in range(2)]
    for r, d in enumerate(res):
        assert d == dict(local_rank=r, rank=r, world=2, sum=3.0, devices="0-1", addr="127.0.0.1")


def test_launch_stops_the_other_ranks_on_failure(tmp_path):
    script = tmp_path / "fail.py"
    script.write_text("import sys, time\nr = int(sys.argv[sys.argv.index('--local_rank') + 1])\n"
                      "sys.exit(3) if r == 1 else time.sleep(120)\n")
    t0 = time.time()
    rc = launch.launch(str(script), [], nproc=2, port=29534)
    assert rc == 3 and time.time() - t0 < 60


def test_main_parses_the_script_tail(tmp_path, monkeypatch):
    seen = {}
    monkeypatch.setattr(launch, "launch", lambda s, args, n, port: seen.update(s=s, args=args, n=n, port=port) or 0)
    assert launch.main(["--nproc", "4", "--port", "29999", "train.py", "-d", "0-3", "-c", "x.pth"]) == 0
    assert seen == dict(s="train.py", args=["-d", "0-3", "-c", "x.pth"], n=4, port=29999)
