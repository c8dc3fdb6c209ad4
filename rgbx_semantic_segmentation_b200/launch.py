"""One-node launcher for the reference's UNCHANGED train.py on torch >= 2 (one process per GPU).

Why it exists: the reference's `Engine` (engine/engine.py:40-58) takes the local rank from the argv flag
`--local_rank` (underscore, engine.py:70), the world size from `WORLD_SIZE` and the rendezvous port from `-p/--port`
(it overwrites `MASTER_PORT` with it, engine.py:55).  `python -m torch.distributed.launch` of torch >= 2 passes
`--local-rank=<r>` (dashed), which that parser rejects, and `torchrun` passes nothing, so every rank would pick GPU 0.
This launcher gives each rank exactly what the reference reads:

    python -m rgbx_semantic_segmentation_b200.launch --nproc 8 [--port 29500] train.py [train.py args...]

    rank r:  RANK=r LOCAL_RANK=r WORLD_SIZE=N MASTER_ADDR=127.0.0.1 MASTER_PORT=<port>
             python train.py --local_rank r -p <port> [train.py args...]

It is plumbing only (no CUDA, no collectives): the first rank that exits non-zero makes the launcher terminate the
others (by their exact PIDs) and return that exit code; SIGINT/SIGTERM are forwarded.
"""
import argparse
import os
import signal
import subprocess
import sys
import threading
import time


def rank_env(rank, nproc, port, base=None, addr="127.0.0.1"):
    """environment of rank `rank` (the keys `init_method='env://'` and the reference's Engine read)"""
    env = dict(os.environ if base is None else base)
    env.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(nproc), MASTER_ADDR=addr, MASTER_PORT=str(port))
    return env


def rank_argv(script, script_args, rank, port, python=None):
    """command line of rank `rank`: the reference parser's spelling of the rank flag and its own port flag"""
    return [python or sys.executable, "-u", script, "--local_rank", str(rank), "-p", str(port), *script_args]


def launch(script, script_args=(), nproc=1, port=29500, python=None, poll_s=0.2, env=None):
    """start `nproc` ranks of `script`, wait for all; returns the first non-zero exit code (0 if every rank succeeded)"""
    if nproc < 1:
        raise ValueError("--nproc must be >= 1")
    procs = []
    try:
        for r in range(nproc):
            procs.append(subprocess.Popen(rank_argv(script, list(script_args), r, port, python),
                                          env=rank_env(r, nproc, port, env)))

        def forward(sig, _frame):
            for p in procs:
                if p.poll() is None:
                    p.send_signal(sig)

        # signal handlers can only be installed from the main thread; elsewhere the ranks are still stopped by the finally below
        main = threading.current_thread() is threading.main_thread()
        old = {s: signal.signal(s, forward) for s in (signal.SIGINT, signal.SIGTERM)} if main else {}
        try:
            rc = 0
            alive = set(range(nproc))
            while alive and rc == 0:
                for r in sorted(alive):
                    code = procs[r].poll()
                    if code is None:
                        continue
                    alive.discard(r)
                    if code != 0:
                        rc = code
                        sys.stderr.write("[cmx_b200.launch] rank %d exited with code %d; stopping the other ranks\n" % (r, code))
                        break
                if alive and rc == 0:
                    time.sleep(poll_s)
            return rc
        finally:
            for s, h in old.items():
                signal.signal(s, h)
    finally:
        for p in procs:
            if p.poll() is None:
                p.terminate()
        deadline = time.time() + 10
        for p in procs:
            try:
                p.wait(timeout=max(0.1, deadline - time.time()))
            except subprocess.TimeoutExpired:
                p.kill()
                p.wait()


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m rgbx_semantic_segmentation_b200.launch", description=__doc__.split("\n\n")[0])
    ap.add_argument("--nproc", type=int, default=None, help="ranks = GPUs on this node (default: all visible GPUs)")
    ap.add_argument("--port", type=int, default=29500, help="rendezvous port (passed to the script as -p and as MASTER_PORT)")
    ap.add_argument("script", help="the reference's train.py (or any script using its Engine)")
    ap.add_argument("script_args", nargs=argparse.REMAINDER)
    a = ap.parse_args(argv)
    nproc = a.nproc
    if nproc is None:
        import torch
        nproc = torch.cuda.device_count()
        if nproc < 1:
            ap.error("no CUDA device visible; pass --nproc explicitly")
    return launch(a.script, a.script_args, nproc, a.port)


if __name__ == "__main__":
    sys.exit(main())
