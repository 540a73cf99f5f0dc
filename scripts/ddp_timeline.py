#!/usr/bin/env python
"""Per-stream CUDA-event timeline of ONE data-parallel step (eager launches, no graph), rank 0's view:
   torchrun --nproc-per-node N scripts/ddp_timeline.py out.json [--nccl-max-ctas K]
Records on the main stream: end of forward, end of every layer's backward, end of backward (all wgrads joined), after
the wait for the last gradient bucket, after the norm, after AdamW.  The same script at N=1 gives the no-communication
reference; the differences say where a multi-GPU step loses its time (slower backward = SM / HBM contention with NCCL's
kernels; a long wait after backward = exposed tail of the all-reduce)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import ecg_b200
from bench import BASE_CFG


def main():
    out = sys.argv[1]
    max_ctas = int(sys.argv[sys.argv.index('--nccl-max-ctas') + 1]) if '--nccl-max-ctas' in sys.argv else None
    world, rank, local = int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        if max_ctas:
            opts = dist.ProcessGroupNCCL.Options()
            opts.config.max_ctas = max_ctas
            dist.init_process_group('nccl', device_id=dev, pg_options=opts)
        else:
            dist.init_process_group('nccl', device_id=dev)
    torch.manual_seed(77)
    model = ecg_b200.EcgVit(config=ecg_b200.EcgVitConfig(compute_dtype='bf16', **BASE_CFG)).to(dev).train()
    tr = ecg_b200.FusedTrainer(model, use_cuda_graph=False)
    x, y = ecg_b200.synthetic_batch(256, length=2500, seed=77 + rank)
    x, y = x.to(dev), y.to(dev)
    for _ in range(5):
        tr.step(x, y)
    torch.cuda.synchronize()
    marks = []

    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        marks.append((name, e))

    eng = model._engine
    hook = model._after_layer_backward

    def layer_done(l):
        if hook is not None:
            hook(l)
        if torch.cuda.current_stream() != eng.side_stream:
            pass

    # wrap the pieces of FusedTrainer._device_step
    fwd, bwd = eng.forward, eng.backward

    def fwd_m(*a, **k):
        mark('step start')
        r = fwd(*a, **k)
        mark('forward done')
        return r

    def bwd_m(*a, **k):
        r = bwd(*a, **k)
        mark('backward done (wgrads joined)')
        return r

    eng.forward, eng.backward = fwd_m, bwd_m
    if tr._reducer is not None:
        fin = tr._reducer.finish

        def fin_m():
            fin()
            mark('last gradient bucket reduced')
        tr._reducer.finish = fin_m
    res = []
    for it in range(10):
        marks.clear()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        tr.step(x, y)
        mark('norm + AdamW done')
        torch.cuda.synchronize()
        t0 = marks[0][1]
        res.append({n: t0.elapsed_time(e) for n, e in marks})
    med = {k: sorted(r[k] for r in res)[len(res) // 2] for k in res[0]}
    if rank == 0:
        json.dump({'world': world, 'nccl_max_ctas': max_ctas, 'eager_ms_median_of_10': med}, open(out, 'w'), indent=1)
        print(json.dumps(med))
    if world > 1:
        torch.cuda.synchronize()
        os._exit(0)


if __name__ == '__main__':
    main()
