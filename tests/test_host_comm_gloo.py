"""eqd_set_host_comm on CPU: two processes (torch.distributed, gloo, world_size 2), one sub-domain each, hand the
library their own all-gather; eqd_sum_shared must then produce the reference's init-time sums over the rank face
(nodalMassArr and fnms through MPI4NodalQuant, assembleGlobalMass.f90:40-41; arn through MPI4arn,
meshgen.f90:274-395) exactly as the host-side restatement does for both sub-domains in one process.  The library's
host code runs against the CUDA-runtime stand-in of tools/hoststub (no kernels run; the sums are host code), in
subprocesses so that the stand-in never shares a process with the real library.  The peer-memory step exchange
itself needs GPUs: tests/test_gpu_nccl.py."""
import os
import socket
import subprocess
import sys
import textwrap

import parity

ROOT = parity.ROOT

SCRIPT = textwrap.dedent(r"""
    import ctypes as C, os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    rank, port, decomp = int(sys.argv[1]), sys.argv[2], tuple(int(x) for x in sys.argv[3].split("x"))
    sys.path.insert(0, %(root)r)
    sys.path.insert(0, os.path.join(%(root)r, "tools", "hoststub"))
    import setup_probe
    setup_probe.OUT = sys.argv[4]
    lib = os.path.join(setup_probe.OUT, "libeqdyna_b200_hoststub.so")
    from eqdyna_b200 import build, cases, device as dev
    from eqdyna_b200.host import World
    build.cuda_lib_path = lambda: lib
    # torch has the real CUDA runtime in the global scope of this process: bind the library to its own dependency (the
    # stand-in) first
    C.CDLL(lib, mode=os.RTLD_NOW | os.RTLD_DEEPBIND)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = port
    dist.init_process_group("gloo", rank=rank, world_size=2)
    calls = []
    def allgather(send, n):
        s = torch.frombuffer(send, dtype=torch.uint8)
        out = torch.empty(n * s.numel(), dtype=torch.uint8)
        dist.all_gather_into_tensor(out, s)
        calls.append(s.numel())
        return out.numpy()
    w = World(cases.materialize("test.tpv8"), np_xyz=decomp, nstep=4)
    w.build(sum_shared=False)        # both sub-domains: the expected sums need the neighbour's
    v = w.view(rank)
    d = dev.Domain(v, host_comm=(2, rank, allgather))
    d.sum_shared()
    assert d.halo_mode() == 1        # the stand-in has no peers to map: the transport falls to NCCL, which this run never needs
    mass = d.fetch(dev.F_MASS, (v.Neq,))
    fnms = d.fetch(dev.F_FNMS, (v.Nn,))
    arn = d.fetch(dev.F_ARN, v.arn.shape)
    try:
        d.run(1, 1)
        raise SystemExit("a run without any transport must fail")
    except dev.StepError as e:
        assert "eqd_set_comm" in str(e), str(e)
    d.close()
    before = (v.nodalMassArr.copy(), v.fnms.copy(), v.arn.copy())
    w.sum_shared()
    assert not np.array_equal(before[0], v.nodalMassArr) and not np.array_equal(before[2], v.arn)   # the face carries mass and fault area
    free = v.eqNumIndexArr[v.eqNumIndexArr > 0] - 1   # equations of non-fixed dofs (fixed nodes keep no mass on the device)
    assert np.array_equal(mass[free], v.nodalMassArr[free])
    assert np.array_equal(fnms, v.fnms)
    assert np.array_equal(arn, v.arn)
    assert len(calls) >= 4, calls     # IPC records + outcome + lengths + vectors of the x phase
    dist.barrier()
    dist.destroy_process_group()
    print("HOSTCOMM_OK", rank, calls)
""")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_sum_shared_over_the_hosts_allgather(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tools", "hoststub"))
    import setup_probe
    setup_probe.OUT = str(tmp_path)
    setup_probe.build_stub()
    port = str(_free_port())
    script = SCRIPT % {"root": ROOT}
    procs = [subprocess.Popen([sys.executable, "-c", script, str(r), port, "2x1x1", str(tmp_path)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "HOSTCOMM_OK %d" % r in out, out[-3000:]


ERRORS = textwrap.dedent(r"""
    import ctypes as C, os, sys
    import numpy as np
    sys.path.insert(0, %(root)r)
    sys.path.insert(0, os.path.join(%(root)r, "tools", "hoststub"))
    import setup_probe
    setup_probe.OUT = sys.argv[1]
    lib = os.path.join(setup_probe.OUT, "libeqdyna_b200_hoststub.so")
    from eqdyna_b200 import build, cases, device as dev
    from eqdyna_b200.host import World
    build.cuda_lib_path = lambda: lib
    w = World(cases.materialize("test.tpv8"), np_xyz=(2, 1, 1), nstep=4)
    w.build(rank=0, sum_shared=False)
    v = w.view(0)
    def broken(send, n):
        raise RuntimeError("the host's communicator is down")
    # rank / nranks must be the sub-domain's own (error 4, as eqd_set_comm)
    for bad in ((3, 0), (2, 1), (2, -1)):
        d = dev.Domain(v)
        try:
            d.set_host_comm(bad[0], bad[1], broken)
            raise SystemExit("accepted nranks, rank = %%r" %% (bad,))
        except dev.StepError as e:
            assert e.code == 4, e
        d.close()
    # a failing all-gather surfaces as an error of the call that needed it, not as a crash
    d = dev.Domain(v, host_comm=(2, 0, broken))
    try:
        d.sum_shared()
        raise SystemExit("eqd_sum_shared went through a broken communicator")
    except dev.StepError as e:
        assert e.code == 3 and "all-gather" in str(e), e
    d.close()
    # after the first eqd_sum_shared / eqd_run the communicator is fixed
    w1 = World(cases.materialize("test.tpv8"), np_xyz=(1, 1, 1), nstep=4)
    w1.build(rank=0, sum_shared=False)
    d = dev.Domain(w1.view(0))
    d.sum_shared()
    try:
        d.set_host_comm(1, 0, broken)
        raise SystemExit("eqd_set_host_comm accepted after finalize")
    except dev.StepError as e:
        assert e.code == 4, e
    d.close()
    print("HOSTCOMM_ERRORS_OK")
""")


def test_host_comm_argument_and_failure_paths(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tools", "hoststub"))
    import setup_probe
    setup_probe.OUT = str(tmp_path)
    setup_probe.build_stub()
    r = subprocess.run([sys.executable, "-c", ERRORS % {"root": ROOT}, str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=600)
    assert r.returncode == 0 and "HOSTCOMM_ERRORS_OK" in r.stdout, r.stdout[-3000:]
