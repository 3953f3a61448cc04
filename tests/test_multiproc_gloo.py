"""N > 1 host logic on CPU: two processes (torch.distributed, gloo, world_size 2),
one sub-domain each, exchange nodal-force halos in the reference's x -> y -> z
order with send/recv of the packed face buffers -- the same packing order and the
same neighbour arithmetic the CUDA library uses with ncclSend/ncclRecv -- and must
reproduce the in-process two-sub-domain run bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parity

NSTEP = 12
CASE = "test.tpv8"
DECOMP = (2, 1, 1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, port, out):
    sys.path.insert(0, parity.ROOT)
    sys.path.insert(0, os.path.join(parity.ROOT, "tests"))
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=2)
    # every process generates both sub-domains (the init-time mass / fnms / arn sums need the
    # neighbour) but steps only its own
    w = parity.build_world(CASE, DECOMP, NSTEP)
    st = oracle.RankStepper(w, rank)
    p = w.view(rank).params
    npx = (p.npx, p.npy, p.npz)
    me = (p.me // (p.npy * p.npz), (p.me // p.npz) % p.npy, p.me % p.npz)
    stride = (p.npy * p.npz, p.npz, 1)
    for nt in range(1, NSTEP + 1):
        st.pre(nt)
        for a in range(3):
            if npx[a] <= 1:
                continue
            sends = {side: st.pack(a, side) for side in (0, 1)}   # both faces hold pre-phase values
            for side in (0, 1):
                active = me[a] != 0 if side == 0 else me[a] != npx[a] - 1
                if not active:
                    continue
                nb = p.me + (-stride[a] if side == 0 else stride[a])
                sbuf = torch.from_numpy(sends[side])
                rbuf = torch.zeros_like(sbuf)
                if p.me < nb:
                    dist.send(sbuf, nb); dist.recv(rbuf, nb)
                else:
                    dist.recv(rbuf, nb); dist.send(sbuf, nb)
                st.add(a, side, rbuf.numpy())
        st.post(nt)
    v = w.view(rank)
    np.savez(os.path.join(out, "rank%d.npz" % rank), disp=v.dispArr, vel=v.velArr, fric=v.fric, acc=v.nodalForceArr)
    dist.barrier()
    dist.destroy_process_group()


def test_two_process_halo_exchange_matches_in_process(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=2, join=True)
    w = parity.build_world(CASE, DECOMP, NSTEP)
    parity.run_oracle(w)
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        v = w.view(r)
        assert np.array_equal(got["disp"], v.dispArr)
        assert np.array_equal(got["vel"], v.velArr)
        assert np.array_equal(got["acc"], v.nodalForceArr)
        assert np.array_equal(got["fric"], v.fric)
    w.close()
