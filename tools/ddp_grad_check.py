#!/usr/bin/env python
"""N-rank data-parallel gradients == single-process gradients of the concatenated batch.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_grad_check.py [--graph]

Every rank runs the native training step on its shard of a fixed batch with model.grad_allreduce = True (bucketed NCCL
all-reduce inside backward, the non-backbone bucket overlapped with the backbone backward); rank 0 then repeats the step on
the whole batch in one process and compares: world * mean-reduced gradient == full-batch gradient of the summed loss.
Prints DDP_GRAD_OK and exits 0 on success."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from sound_event_detection_transformer_b200 import parallel, spec, synth
from sound_event_detection_transformer_b200.sedt import build_model


def loss_fn(out, R, sl):
    tot = (out["pred_logits"] * R["l"][-1][sl]).sum() + (out["pred_boxes"] * R["b"][-1][sl]).sum() + (out["at"] * R["a"][sl]).sum()
    for i, aux in enumerate(out["aux_outputs"]):
        tot = tot + (aux["pred_logits"] * R["l"][i][sl]).sum() + (aux["pred_boxes"] * R["b"][i][sl]).sum()
    return tot


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--overlap", action="store_true", help="bucketed all-reduce overlapped with the backbone backward")
    ap.add_argument("--per-rank", type=int, default=2)
    o = ap.parse_args()
    rank, world, local = parallel.env_ranks()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    parallel.init_from_env("nccl", dev)
    args = spec.config_args("c1")
    args.enc_layers, args.dec_layers, args.dropout, args.precision = 2, 2, 0.0, "bf16"
    sd = synth.synth_state_dict(args, 21)
    B = o.per_rank * world
    clips = synth.synth_clips(B, 200, 64, seed=30)
    g = torch.Generator().manual_seed(22)
    D, Q, C = args.dec_layers, args.num_queries, args.num_classes
    R = {"l": torch.randn(D, B, Q, C + 1, generator=g).to(dev), "b": torch.randn(D, B, Q, 2, generator=g).to(dev),
         "a": torch.randn(B, C, generator=g).to(dev)}

    def grads(sl, allreduce):
        model, _, _ = build_model(args)
        model.load_state_dict(sd, strict=True)
        model = model.to(dev).train()
        model.grad_allreduce, model.use_cuda_graph, model.grad_allreduce_overlap = allreduce, o.graph, o.overlap
        out = None
        for _ in range(3 if o.graph else 1):                 # graph mode: capture, then replays
            model.zero_grad(set_to_none=True)
            out = model(clips[sl].to(dev))
            loss_fn(out, R, sl).backward()
        torch.cuda.synchronize()
        return {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.requires_grad}

    lo, hi = parallel.shard_range(B, rank, world)
    shard = grads(slice(lo, hi), True)
    parallel.barrier()
    ok = True
    if rank == 0:
        full = grads(slice(0, B), False)
        worst = (0.0, "")
        for n, gf in full.items():
            got = shard[n] * world
            rel = ((got - gf).norm() / gf.norm().clamp_min(1e-20)).item()
            if float(gf.norm()) == 0.0 and float(got.norm()) == 0.0:
                continue
            worst = max(worst, (rel, n))
        ok = worst[0] < 2e-3
        print(f"{'DDP_GRAD_OK' if ok else 'DDP_GRAD_MISMATCH'} world={world} graph={o.graph} tensors={len(full)} "
              f"worst rel-L2 {worst[0]:.2e} ({worst[1]})", flush=True)
    parallel.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
