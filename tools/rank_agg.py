"""Cross-rank aggregation used by bench.py: times are the MAX over ranks, work is the SUM over ranks.  In the default
sharded mode every rank holds a slice of the same bond update (its own share of the apply flops); with
`--multi replicas` the ranks run independent copies."""
import torch
import torch.distributed as dist


def aggregate(ms_per_step: float, solver_s: float, work_flops: float, device="cuda"):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(ms_per_step=ms_per_step, solver_s=solver_s, work_flops=work_flops,
                    tflops=work_flops / solver_s / 1e12 if solver_s > 0 else 0.0)
    t = torch.tensor([ms_per_step, solver_s], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    w = torch.tensor([work_flops], dtype=torch.float64, device=device)
    dist.all_reduce(w, op=dist.ReduceOp.SUM)
    return dict(ms_per_step=t[0].item(), solver_s=t[1].item(), work_flops=w[0].item(),
                tflops=w[0].item() / t[1].item() / 1e12 if t[1].item() > 0 else 0.0)
