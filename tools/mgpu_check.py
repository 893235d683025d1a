"""Multi-GPU parity check (run under torchrun): the echelon form, the RREF and the kernel basis computed cooperatively
by all ranks must be identical, array for array, to the ones each rank computes alone, and the cooperative run must
have moved bytes through NCCL (tests/test_gpu_multi_rank.py asserts on the last line)."""
import os, sys, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
os.environ["SPASM_B200_DEVICE"] = str(local)
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import spasm_b200, oracle, util
from spasm_b200 import synthetic, host, sharding
L = spasm_b200.lib(); L.spasm_b200_set_verbose(0)
cases = [synthetic.config1(0.1), synthetic.config2(0.1).transposed(), synthetic.config5(0.05), synthetic.config3(0.02)]
opts = [{}, {}, {}, {"sparsity_threshold": 0.01}]
cases.append(synthetic.config4(0.02)); opts.append({})
def digest(f):
    U = f.U
    h = hashlib.sha256()
    for k in "pjx": h.update(np.ascontiguousarray(U[k]).tobytes())
    h.update(f.qinv.tobytes())
    Rm, Rqinv = host.rref(L, f)
    Km = host.kernel(L, f)
    for M in (Rm.numpy(), Km.numpy()):
        for k in "pjx": h.update(np.ascontiguousarray(M[k]).tobytes())
    h.update(np.ascontiguousarray(Rqinv).tobytes())
    digest.rref_nnz = Rm.nnz
    return f.rank, h.hexdigest()
alone = []
for t, o in zip(cases, opts):
    A = host.compress(L, t); oracle.reset_rand()
    alone.append(digest(host.echelonize(L, A, host.default_opts(L, **o))))
sharding.init_comm(L, dist, device=torch.device("cuda", local))
ok = True
total_bytes = 0
for (t, o), want in zip(zip(cases, opts), alone):
    A = host.compress(L, t); oracle.reset_rand(); L.spasm_b200_reset_stats()
    got = digest(host.echelonize(L, A, host.default_opts(L, **o)))
    s = util.product_stats(L)
    same = got == want
    ok &= same and s.nccl_bytes > 0
    total_bytes += s.nccl_bytes
    print(f"rank {dist.get_rank()}/{dist.get_world_size()} {t.name}: rank {got[0]} identical={same} nccl_bytes={s.nccl_bytes}", flush=True)
# gather-to-root mode (spasm_b200_comm_result_root): rank 0 holds the same result, the other ranks an empty matrix
L.spasm_b200_comm_result_root(0)
for (t, o), want in list(zip(zip(cases, opts), alone))[-2:]:
    A = host.compress(L, t); oracle.reset_rand(); L.spasm_b200_reset_stats()
    got = digest(host.echelonize(L, A, host.default_opts(L, **o)))
    same = (got == want) if dist.get_rank() == 0 else (got[0] == want[0] and digest.rref_nnz == 0)
    ok &= same
    print(f"rank {dist.get_rank()}/{dist.get_world_size()} {t.name} (result on rank 0): identical={same} rref_nnz={digest.rref_nnz}", flush=True)
L.spasm_b200_comm_result_root(-1)
v = torch.tensor([1.0 if ok else 0.0], device="cuda"); dist.all_reduce(v, op=dist.ReduceOp.MIN)
if dist.get_rank() == 0: print("MULTI-GPU PARITY", "OK" if v.item() == 1.0 else "FAILED", f"world={dist.get_world_size()} nccl_bytes_rank0={total_bytes}", flush=True)
L.spasm_b200_comm_destroy(); dist.destroy_process_group()
