"""
Multi-GPU from a single-process call (SURVEY.md 8b "owns its CUDA streams/NCCL comm internally", VERDICT r1
missing #3).  The reference calls ``run_plmc`` from ONE blocking Python process and forwards ``cpu`` as plmc's
``-n`` (evcouplings/couplings/tools.py:257-259); to give that call all GPUs of the box this module starts one
rank per GPU (``python -m evcouplings_b200.worker``), each of which runs the same ``tools.run_plmc`` as a member
of a torch.distributed / NCCL group on 127.0.0.1: sequences sharded over ranks, ONE all-reduce of [g, -loglk]
per evaluation, rank 0 writes the files.  The parent waits, relays failures as ExternalToolError and rebuilds
the PlmcResult from rank 0's plmc-style log (same parser as the reference's, tools.py:20-108).
"""
import json
import os
import pickle
import socket
import subprocess
import sys
import tempfile
import time


def _free_port():
    s = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def run_plmc_multi_gpu(ndev, kwargs, return_run=False, backend="nccl", engine_factory=None, timeout=None):
    """Run tools.run_plmc(**kwargs) on ``ndev`` ranks (one per GPU).  ``engine_factory`` ("module:attr") and
    ``backend`` exist for the CPU/gloo test of this plumbing; the product default is the CUDA engine over NCCL."""
    from . import tools
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    workdir = tempfile.mkdtemp(prefix="evcplm_ranks_")
    spec_path = os.path.join(workdir, "spec.json")
    result_path = os.path.join(workdir, "rank0.pkl")
    with open(spec_path, "w") as f:
        json.dump(dict(kwargs=kwargs, backend=backend, engine_factory=engine_factory, result=result_path), f)
    port = _free_port()
    procs, logs = [], []
    for r in range(ndev):
        env = dict(os.environ)
        env.update(RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(ndev), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
        env.pop("EVC_NUM_GPUS", None)
        # like torchrun: one OpenMP / BLAS thread per rank unless the caller decided otherwise -- ndev ranks each
        # spinning a full-size host thread pool starve the threads that drive the GPUs
        for var in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
            env.setdefault(var, "1")
        log = open(os.path.join(workdir, "rank%d.err" % r), "w+")
        logs.append(log)
        procs.append(subprocess.Popen([sys.executable, "-m", "evcouplings_b200.worker", spec_path],
                                      env=env, stdout=log, stderr=subprocess.STDOUT, cwd=root))
    t0 = time.time()
    failed = None
    try:
        while True:
            codes = [p.poll() for p in procs]
            bad = [r for r, c in enumerate(codes) if c not in (None, 0)]
            if bad:
                failed = bad[0]
                break
            if all(c == 0 for c in codes):
                break
            if timeout is not None and time.time() - t0 > timeout:
                failed = -1
                break
            time.sleep(0.05)
    finally:
        for p in procs:                       # our own children, by PID
            if p.poll() is None:
                if failed is not None:
                    p.terminate()
                try:
                    p.wait(timeout=30)
                except subprocess.TimeoutExpired:
                    p.kill()
    if failed is not None:
        tail = ""
        if failed >= 0:
            logs[failed].seek(0)
            tail = logs[failed].read()[-2000:]
        for log in logs:
            log.close()
        raise tools.ExternalToolError("multi-GPU plmc run failed (%s): %s"
                                      % ("timeout" if failed < 0 else "rank %d" % failed, tail))
    if os.environ.get("EVC_TRACE"):
        for r, log in enumerate(logs):
            log.seek(0)
            for ln in log.read().splitlines():
                if "evc-trace" in ln:
                    sys.stderr.write("[rank %d] %s\n" % (r, ln))
    for log in logs:
        log.close()
    with open(result_path, "rb") as f:
        rec = pickle.load(f)
    run = tools.PlmcRun()
    from . import lbfgs as _lbfgs
    run.log, run.timings, run.n_eff = rec["log"], rec["timings"], rec["n_eff"]
    run.lbfgs = _lbfgs.LbfgsResult(*rec["lbfgs"]) if rec["lbfgs"] is not None else None
    run.timings["ranks"] = ndev
    run.timings["launcher_total_s"] = time.time() - t0
    iter_df, fields = tools.parse_plmc_log(run.log)
    k = kwargs
    tools._require_file("plmc returned no couplings", k["couplings_file"])
    if k.get("param_file") is not None:
        tools._require_file("plmc returned no parameter file", k["param_file"])
    result = tools.PlmcResult(k["couplings_file"], k.get("param_file"), iter_df, *fields)
    run.result = result
    for name in os.listdir(workdir):
        os.unlink(os.path.join(workdir, name))
    os.rmdir(workdir)
    if return_run:
        return result, run
    return result
