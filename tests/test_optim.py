"""Optimiser step and gradient all-reduce of data-parallel training.

Oracle: oracle/adam_oracle.py, a numpy restatement of torch's single-tensor Adam — the algorithm behind the reference's
`torch.optim.Adam(l, lr=0.0, eps=1e-15)` (gaussian_model.py:201) — pinned here against torch.optim.Adam itself on the CPU.
Tolerance: 2e-6 of the largest entry per tensor after 6 steps (fp32; torch's CPU kernels may contract a multiply-add or turn a
division by a scalar into a multiplication, 1 ulp each; the CUDA kernel follows the restatement operation for operation)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, grad_close
from oracle.adam_oracle import adam_step

TOL = 2e-6
ORC_TOL = 5e-7        # CUDA kernel vs the numpy restatement: same operations; the host-side bias corrections may round differently by an ulp
GROUPS = [("xyz", (1000, 3), 1.6e-4), ("f_dc", (1000, 1, 3), 2.5e-3), ("f_rest", (1000, 15, 3), 1.25e-4), ("opacity", (1000, 1), 0.05),
          ("scaling", (1000, 2), 5e-3), ("rotation", (1000, 4), 1e-3)]       # configs/exp.yaml learning rates


def _problem(seed, device="cpu", groups=GROUPS):
    g = torch.Generator().manual_seed(seed)
    params = [torch.randn(shape, generator=g).to(device).requires_grad_(True) for _, shape, _ in groups]
    grads = [[(torch.randn(shape, generator=g) * (10.0 ** float(torch.randint(-4, 2, (1,), generator=g)))).to(device) for _, shape, _ in groups]
             for _ in range(6)]
    return params, grads


def _oracle_run(params, grads, groups=GROUPS, eps=1e-15):
    P = [p.detach().cpu().numpy().copy() for p in params]
    M = [np.zeros_like(p) for p in P]; V = [np.zeros_like(p) for p in P]
    for step, gs in enumerate(grads, 1):
        for i, (_, _, lr) in enumerate(groups):
            P[i], M[i], V[i] = adam_step(P[i], gs[i].cpu().numpy(), M[i], V[i], lr, step, eps=eps)
    return P, M, V


def test_oracle_vs_torch_adam_cpu():
    params, grads = _problem(0)
    opt = torch.optim.Adam([{"params": [p], "lr": lr, "name": n} for p, (n, _, lr) in zip(params, GROUPS)], lr=0.0, eps=1e-15, foreach=False)
    ref = [p.detach().clone() for p in params]
    for gs in grads:
        for p, g in zip(params, gs):
            p.grad = g.clone()
        opt.step()
    P, M, V = _oracle_run(ref, grads)
    for i, (n, _, _) in enumerate(GROUPS):
        grad_close(P[i], params[i].detach().numpy(), TOL, f"param {n}")
        grad_close(M[i], opt.state[params[i]]["exp_avg"].numpy(), TOL, f"exp_avg {n}")
        grad_close(V[i], opt.state[params[i]]["exp_avg_sq"].numpy(), TOL, f"exp_avg_sq {n}")
        assert not np.allclose(P[i], ref[i].numpy()), "the step must move the parameters"


def test_fused_adam_has_torch_adams_interface_and_fails_loudly_on_cpu():
    from lidar_rt_b200 import native
    from lidar_rt_b200.optim import FusedAdam
    p = torch.zeros(4, 3, requires_grad=True)
    opt = FusedAdam([{"params": [p], "lr": 0.1, "name": "xyz"}], lr=0.0, eps=1e-15)
    assert opt.param_groups[0]["name"] == "xyz" and opt.param_groups[0]["lr"] == 0.1 and opt.param_groups[0]["eps"] == 1e-15
    assert opt.param_groups[0]["betas"] == (0.9, 0.999)
    opt.step()                                    # no gradients: nothing to do, no native call
    p.grad = torch.ones_like(p)
    with pytest.raises(native.LrtError):
        opt.step()                                # CPU parameters: there is no fallback
    sd = opt.state_dict(); opt.load_state_dict(sd)


def test_fused_adam_state_dict_is_plain_and_picklable():
    import pickle
    from lidar_rt_b200.optim import FusedAdam, _State
    p = torch.zeros(4, 3, requires_grad=True)
    opt = FusedAdam([{"params": [p], "lr": 0.1, "name": "xyz"}], lr=0.0, eps=1e-15)
    st = _State(step=torch.tensor(3.0), exp_avg=torch.ones(4, 3), exp_avg_sq=torch.ones(4, 3)); st.owner = opt
    opt.state[p] = st
    sd = opt.state_dict()
    assert type(sd["state"][0]) is dict and set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    blob = pickle.dumps(sd)
    assert len(blob) < 4000, "a saved state must not drag the optimizer along"
    opt2 = FusedAdam([{"params": [torch.zeros(4, 3, requires_grad=True)], "lr": 0.1, "name": "xyz"}], lr=0.0, eps=1e-15)
    opt2.load_state_dict(pickle.loads(blob))
    assert int(opt2.state_dict()["state"][0]["step"]) == 3
    st["exp_avg"] = torch.zeros(4, 3)               # an outside edit (densification) marks the native table stale
    assert opt._dirty


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "lidar-rt_b200"))
from lidar_rt_b200.optim import all_reduce_gradients
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
shapes = [(5, 3), (5, 15, 3), (5, 1), (7, 4)]
ps = [torch.zeros(s, requires_grad=True) for s in shapes]
g = torch.Generator().manual_seed(100 + rank)
for i, p in enumerate(ps):
    if not (rank == 1 and i == 2):              # rank 1 has no gradient for parameter 2: counts as zero
        p.grad = torch.randn(p.shape, generator=g)
flat = all_reduce_gradients(ps)
exp = []
for i, s in enumerate(shapes):
    acc = torch.zeros(s)
    for r in range(world):
        gr = torch.Generator().manual_seed(100 + r)
        for j, sj in enumerate(shapes):
            if not (r == 1 and j == 2):
                t = torch.randn(sj, generator=gr)
                if j == i:
                    acc += t
    exp.append(acc / world)
ok = all(torch.allclose(p.grad, e, atol=1e-6) for p, e in zip(ps, exp))
ok = ok and flat.numel() == sum(p.numel() for p in ps) and all(p.grad.data_ptr() >= flat.data_ptr() for p in ps)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
"""


def test_packed_gradient_all_reduce_world2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    port = 29650 + os.getpid() % 200
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env))
    assert [p.wait(timeout=120) for p in procs] == [0, 0]


def test_packed_gradient_single_process_is_identity():
    from lidar_rt_b200.optim import all_reduce_gradients
    ps = [torch.zeros(3, 2, requires_grad=True), torch.zeros(4, requires_grad=True)]
    ps[0].grad = torch.arange(6.0).reshape(3, 2)
    flat = all_reduce_gradients(ps, world_size=1)
    assert torch.equal(ps[0].grad, torch.arange(6.0).reshape(3, 2)) and torch.equal(ps[1].grad, torch.zeros(4)) and flat.numel() == 10


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_fused_adam_vs_oracle_and_torch_cuda():
    from lidar_rt_b200.optim import FusedAdam
    params, grads = _problem(1, "cuda")
    start = [p.detach().clone() for p in params]
    tparams = [p.detach().clone().requires_grad_(True) for p in params]
    mk = lambda cls, ps, **kw: cls([{"params": [p], "lr": lr, "name": n} for p, (n, _, lr) in zip(ps, GROUPS)], lr=0.0, eps=1e-15, **kw)
    opt, topt = mk(FusedAdam, params), mk(torch.optim.Adam, tparams, foreach=False)
    for gs in grads:
        for p, tp, g in zip(params, tparams, gs):
            p.grad = g.clone(); tp.grad = g.clone()
        opt.step(); topt.step()
    P, M, V = _oracle_run(start, grads)
    for i, (n, _, _) in enumerate(GROUPS):
        st = opt.state[params[i]]
        assert set(st.keys()) == {"step", "exp_avg", "exp_avg_sq"}      # what the reference's densification code edits
        grad_close(params[i].detach().cpu().numpy(), P[i], ORC_TOL, f"param {n} vs oracle")
        grad_close(st["exp_avg"].cpu().numpy(), M[i], ORC_TOL, f"exp_avg {n} vs oracle")
        grad_close(st["exp_avg_sq"].cpu().numpy(), V[i], ORC_TOL, f"exp_avg_sq {n} vs oracle")
        grad_close(params[i].detach().cpu().numpy(), tparams[i].detach().cpu().numpy(), TOL, f"param {n} vs torch.optim.Adam (CUDA)")
    sd = opt.state_dict()
    assert all(int(s_["step"]) == len(grads) for s_ in sd["state"].values())


@pytest.mark.gpu
def test_step_many_sizes_alignment_and_state_edits():
    """All assets in one launch; tensors of 1 / 7 / 2049 / 1M elements; a parameter that is an unaligned view (scalar path);
    state replaced the way the reference's densification does (fresh zeros, step restarts)."""
    from lidar_rt_b200.optim import FusedAdam, step_many
    g = torch.Generator().manual_seed(5)
    sizes = [1, 7, 2049, 1_000_003]
    backing = [torch.randn(n + 1, generator=g).cuda() for n in sizes]
    params = [b[1:].detach().requires_grad_(True) if i % 2 else b[:-1].detach().clone().requires_grad_(True) for i, b in enumerate(backing)]
    opts = [FusedAdam([{"params": [p], "lr": 1e-2 * (i + 1), "name": f"t{i}"}], lr=0.0, eps=1e-15) for i, p in enumerate(params)]
    P = [p.detach().cpu().numpy().copy() for p in params]; M = [np.zeros_like(x) for x in P]; V = [np.zeros_like(x) for x in P]
    steps = [0] * len(params)
    for it in range(4):
        gs = [torch.randn(p.shape, generator=g).cuda() for p in params]
        for p, gr in zip(params, gs):
            p.grad = gr
        if it == 2:                                  # "densification": optimizer state of tensor 2 re-created
            st = opts[2].state[params[2]]
            st["exp_avg"] = torch.zeros_like(params[2]); st["exp_avg_sq"] = torch.zeros_like(params[2]); st["step"] = torch.tensor(0.0)
            M[2][:] = 0; V[2][:] = 0; steps[2] = 0
        step_many(opts)
        for i in range(len(params)):
            steps[i] += 1
            P[i], M[i], V[i] = adam_step(P[i], gs[i].cpu().numpy(), M[i], V[i], 1e-2 * (i + 1), steps[i], eps=1e-15)
    for i in range(len(params)):
        grad_close(params[i].detach().cpu().numpy(), P[i], ORC_TOL, f"tensor {i}")
        grad_close(opts[i].state[params[i]]["exp_avg_sq"].cpu().numpy(), V[i], ORC_TOL, f"exp_avg_sq {i}")
    assert float(backing[1][0]) == float(backing[1][0]) and torch.isfinite(backing[3]).all()      # neighbours of the views untouched / finite


@pytest.mark.gpu
def test_adam_error_behaviour():
    from lidar_rt_b200 import native
    ctx = native.Context("cuda:0")
    p = torch.zeros(8, device="cuda")
    with pytest.raises(native.LrtError):
        ctx.adam_step([(p, p.clone(), p.clone(), torch.zeros(4, device="cuda"), 0.1, 1)], 0.9, 0.999, 1e-8)
    with pytest.raises(native.LrtError):
        ctx.adam_step([(p, p.double(), p.clone(), p.clone(), 0.1, 1)], 0.9, 0.999, 1e-8)
    with pytest.raises(native.LrtError):
        ctx.adam_step([(p, p.clone(), p.clone(), p.clone(), 0.1, 0)], 0.9, 0.999, 1e-8)      # step must be >= 1
    ctx.adam_step([], 0.9, 0.999, 1e-8)
    ctx.close()
