"""In-process A/B of one launcher knob on the whole training step and the sampling forward.

    python profiles/ab_step.py DDPM_HALO_DYNAMIC 0 1 [--rounds 6] [--steps 10] [--no-sampling]

Two boxes differ by +-2 % (power state), more than most single changes, so the two settings are measured in ONE process:
the knob is read by the C library at launch time, i.e. when a CUDA graph is CAPTURED, so one graph is captured per value
and the two graphs are replayed alternately (A B A B ...), `steps` replays per turn, CUDA events around each turn.
Prints one JSON line.  Under torchrun (N > 1) the step is the DDP step and times are the max over ranks.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("knob")
    ap.add_argument("a")
    ap.add_argument("b")
    ap.add_argument("--rounds", type=int, default=6)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--sampling-batch", type=int, default=32)
    ap.add_argument("--no-sampling", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    args = ap.parse_args()

    import torch.distributed as dist
    from polyp_image_generator_b200 import DDPMScheduler, FusedAdamW, UNet2DModel
    from polyp_image_generator_b200.graphs import GraphedTrainStep, GraphedUNetForward
    from polyp_image_generator_b200.model import polyp_unet_config

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def alternate(run_a, run_b):
        ta, tb = [], []
        for _ in range(args.rounds):
            for run, acc in ((run_a, ta), (run_b, tb)):
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.steps):
                    run()
                e1.record()
                barrier()
                t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                acc.append(float(t[0]))
        med = lambda v: sorted(v)[len(v) // 2]
        return {"a_ms": [round(x, 3) for x in ta], "b_ms": [round(x, 3) for x in tb],
                "a_median": round(med(ta), 3), "b_median": round(med(tb), 3),
                "b_over_a": round(med(tb) / med(ta), 4)}

    out = {"knob": args.knob, "a": args.a, "b": args.b, "n_gpus": world, "steps_per_turn": args.steps}
    S, B = args.size, args.batch
    if not args.no_train:
        steps = {}
        for val in (args.a, args.b):
            os.environ[args.knob] = val
            torch.manual_seed(0)
            model = UNet2DModel(**polyp_unet_config(S)).to(dev)
            model.train()
            net = model
            if world > 1:
                from polyp_image_generator_b200.ddp import DistributedDataParallel
                net = DistributedDataParallel(model)
            opt = FusedAdamW(model.parameters(), lr=1e-4, max_grad_norm=1.0)
            g = torch.Generator(device=dev).manual_seed(1 + rank)
            clean = torch.rand((B, 3, S, S), device=dev, generator=g) * 2 - 1
            noise = torch.randn((B, 3, S, S), device=dev, generator=g)
            t = torch.randint(0, 1000, (B,), device=dev, dtype=torch.int64, generator=g)
            gs = GraphedTrainStep(net, DDPMScheduler(num_train_timesteps=1000), opt, clean.shape, max_grad_norm=1.0,
                                  warmup_batch=(clean, noise, t))
            steps[val] = (gs, clean, noise, t, model, net, opt)
        os.environ.pop(args.knob, None)
        ra = lambda: steps[args.a][0](*steps[args.a][1:4])
        rb = lambda: steps[args.b][0](*steps[args.b][1:4])
        for _ in range(3):
            ra(), rb()
        out["train_step"] = alternate(ra, rb)
        out["train_step"]["loss_a"] = float(steps[args.a][0].loss)
        out["train_step"]["loss_b"] = float(steps[args.b][0].loss)
        if rank == 0:
            print(json.dumps({"partial": out}), flush=True)
        del steps, ra, rb
        torch.cuda.empty_cache()
    if not args.no_sampling:
        Bs = args.sampling_batch
        fw = {}
        for val in (args.a, args.b):
            os.environ[args.knob] = val
            torch.manual_seed(0)
            model = UNet2DModel(**polyp_unet_config(S)).to(dev)
            model.eval()
            gf = GraphedUNetForward(model, Bs, S, S)
            x = torch.randn((Bs, 3, S, S), device=dev)
            gf(x, 500)
            fw[val] = (gf, x, model)
        os.environ.pop(args.knob, None)
        ra = lambda: fw[args.a][0](fw[args.a][1], 500)
        rb = lambda: fw[args.b][0](fw[args.b][1], 500)
        for _ in range(3):
            ra(), rb()
        out["sampling_forward"] = alternate(ra, rb)
        out["sampling_forward"]["max_abs_diff"] = float((fw[args.a][0].out - fw[args.b][0].out).abs().max())
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        barrier()
        os._exit(0)      # a live graph with captured collectives blocks destroy_process_group


if __name__ == "__main__":
    main()
