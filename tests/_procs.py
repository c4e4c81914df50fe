"""Process plumbing for the multi-rank CPU tests (gloo, world_size > 1): a free rendezvous port per
test instead of fixed numbers (a leftover process of an earlier, interrupted run must not be able
to wedge the suite), fail-fast when a rank dies, and children that are ALWAYS reaped — also when the
test fails or times out."""
import os
import queue
import signal
import socket
import subprocess
import time

import torch.multiprocessing as mp


def free_port():
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_world(target, world, make_args, timeout=600):
    """Spawn `world` ranks of `target(*make_args(rank, port, q))`, return what rank 0 put into `q`.
    Raises AssertionError when a rank exits non-zero, or nothing arrives within `timeout` seconds."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=target, args=tuple(make_args(r, port, q))) for r in range(world)]
    try:
        for p in procs:
            p.start()
        deadline = time.time() + timeout
        while True:
            try:
                result = q.get(timeout=1.0)
                break
            except queue.Empty:
                codes = [p.exitcode for p in procs]
                if any(c not in (None, 0) for c in codes):
                    raise AssertionError(f"a rank died before reporting: exit codes {codes}")
                if all(c == 0 for c in codes):
                    raise AssertionError("all ranks exited without reporting a result")
                if time.time() > deadline:
                    raise AssertionError(f"no result within {timeout} s (exit codes {codes})")
        for p in procs:
            p.join(max(1.0, min(120.0, deadline - time.time() + 60.0)))
        codes = [p.exitcode for p in procs]
        assert all(c == 0 for c in codes), codes
        return result
    finally:
        for p in procs:
            if p.is_alive():
                p.terminate()
        for p in procs:
            p.join(5)
            if p.is_alive():
                p.kill()
                p.join(5)


def run_torchrun(script, nproc, env, timeout, cwd, python, args=()):
    """`python -m torch.distributed.run` on a free port in its own process group; on timeout the
    whole group (agent + workers) is killed, nothing is left behind.  -> (returncode, stdout, stderr)"""
    cmd = [python, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), script, *args]
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=cwd,
                            env=env, start_new_session=True)
    try:
        out, err = proc.communicate(timeout=timeout)
        return proc.returncode, out, err
    except subprocess.TimeoutExpired:
        try:
            os.killpg(proc.pid, signal.SIGKILL)   # the session started above: agent + its workers
        except ProcessLookupError:
            pass
        out, err = proc.communicate()
        return -9, out, err + f"\n[timeout after {timeout} s: process group killed]"
    finally:
        if proc.poll() is None:
            try:
                os.killpg(proc.pid, signal.SIGKILL)
            except ProcessLookupError:
                pass
