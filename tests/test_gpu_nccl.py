"""Multi-process / multi-GPU parity (-m gpu, needs >= 2 devices; skipped on a
single-GPU box): one process per sub-domain, nodal-force halos over peer memory
(or ncclSend/ncclRecv) exactly as `bench.py --gpus N` runs them, compared with the
CPU oracle of the same decomposition at the contract tolerance."""
import os
import socket
import sys

import numpy as np
import pytest

import parity

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, case, decomp, nstep, overlap, own_rank_only, out, march=0):
    sys.path.insert(0, parity.ROOT)
    sys.path.insert(0, os.path.join(parity.ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from eqdyna_b200 import device as dev
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    if own_rank_only:
        # what bench.py --gpus N does: every process builds only its own sub-domain, the operators and
        # the lumped mass are computed on the device and the init-time shared sums go over NCCL
        from eqdyna_b200 import cases
        from eqdyna_b200.host import World
        w = World(cases.materialize(case), np_xyz=decomp, nstep=nstep)
        w.build(rank=rank, sum_shared=False)
    else:
        w = parity.build_world(case, decomp, nstep)   # whole world on the host: init-time shared sums done
    v = w.view(rank)
    d = dev.Domain(v, device=rank % torch.cuda.device_count(), compute_ops=bool(own_rank_only),
                   options={"march": 1} if march else None)
    if march:
        d.set_option("box", 2)
        d.set_option("box_compact", 1)
    if march == 2:
        d.set_option("halo", 0)      # the ncclSend / ncclRecv exchange instead of peer memory
    if march == 3:
        # eqd_set_host_comm: the set-up exchanges go over the host's own all-gather (here gloo, MPI_Allgather in the
        # Fortran host), the steps over peer memory; the library creates no NCCL communicator
        def allgather(send, nr):
            s = torch.frombuffer(send, dtype=torch.uint8)
            out_t = torch.empty(nr * s.numel(), dtype=torch.uint8)
            dist.all_gather_into_tensor(out_t, s)
            return out_t.numpy()
        d.set_host_comm(world_size, rank, allgather)
    else:
        ids = [dev.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        d.set_comm(ids[0], world_size, rank)
    if own_rank_only:
        d.sum_shared()
    d.set_option("overlap", overlap)   # 1: halo on the communication stream under the interior tiles
    n = v.nstep
    d.run(1, n // 2)          # two chunks: the halo state must survive a return to the host
    if march == 3:
        assert d.halo_mode() == 2, "peer memory carries the steps when only the host's communicator is given"
    d.run(n // 2 + 1, n)
    d.fetch_into_view()
    np.savez(os.path.join(out, "rank%d.npz" % rank), disp=v.dispArr, vel=v.velArr, fric=v.fric, fnft=v.fnft,
             hist=v.onFaultQuantHistSCECForm)
    dist.barrier()
    d.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("case,decomp,nstep,overlap,own_rank_only,march",
                         [("test.tpv8", (2, 1, 1), 0, 1, 0, 0), ("test.tpv104", (1, 2, 1), 60, 1, 0, 0), ("test.tpv8", (2, 1, 1), 40, 0, 0, 0),
                          ("test.tpv104", (1, 2, 1), 60, 1, 1, 0), ("test.tpv10", (2, 1, 2), 60, 2, 0, 0),
                          ("test.tpv104", (2, 1, 1), 60, 1, 1, 1), ("test.tpv104", (1, 2, 1), 60, 0, 0, 1), ("test.tpv104", (2, 1, 1), 60, 2, 1, 1),
                          ("test.tpv8", (1, 1, 2), 40, -1, 0, 2), ("test.tpv104", (2, 1, 1), 60, -1, 1, 3), ("test.tpv10", (1, 1, 2), 40, -1, 0, 3)],
                         ids=["tpv8-2x1x1", "tpv104-1x2x1-fault-on-rank-face", "tpv8-2x1x1-no-overlap",
                              "tpv104-1x2x1-own-rank-device-ops-sum-shared", "tpv10-2x1x2-face-tiles-first",
                              "tpv104-2x1x1-march-own-rank-device-ops", "tpv104-1x2x1-march-no-overlap",
                              "tpv104-2x1x1-march-overlap-2-boundary-list-first", "tpv8-1x1x2-march-nccl-sendrecv-instead-of-peer-memory",
                              "tpv104-2x1x1-march-host-communicator-own-rank-sum-shared", "tpv10-1x1x2-march-host-communicator"])
def test_nccl_processes_match_oracle(tmp_path, case, decomp, nstep, overlap, own_rank_only, march):
    import torch
    import torch.multiprocessing as mp
    n = decomp[0] * decomp[1] * decomp[2]
    if torch.cuda.device_count() < n:
        pytest.skip("needs %d GPUs" % n)
    port = _free_port()
    mp.spawn(_worker, args=(n, port, case, decomp, nstep, overlap, own_rank_only, str(tmp_path), march), nprocs=n, join=True)
    wo = parity.build_world(case, decomp, nstep)
    parity.run_oracle(wo)
    for r in range(n):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        o = wo.view(r)
        assert parity.rel_l2(got["disp"], o.dispArr) <= parity.REL_L2_TOL
        assert parity.rel_l2(got["vel"], o.velArr) <= parity.REL_L2_TOL
        k = int(o.nftnd[0])
        if k:
            assert parity.rel_l2(got["fric"][70:80, :k, 0], o.fric[70:80, :k, 0]) <= parity.REL_L2_TOL
            both = (got["fnft"][:k, 0] < 5000.0) & (o.fnft[:k, 0] < 5000.0)
            assert np.array_equal(got["fnft"][:k, 0] < 5000.0, o.fnft[:k, 0] < 5000.0)
            if both.any():
                assert np.max(np.abs(got["fnft"][:k, 0][both] - o.fnft[:k, 0][both])) <= o.params.dt * (1 + 1e-9)
    wo.close()
