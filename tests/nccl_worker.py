"""2-rank NCCL worker of tests/test_parity_full_gpu.py::test_two_rank_nccl_gradients_equal_single_rank.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/nccl_worker.py

Every rank holds the same weights and HALF of a fixed batch.  Checked, eager and under CUDA-graph replay:
  * the data-parallel gradients (mean of the rank gradients, all-reduced over NCCL in buckets on the side stream while
    backward is still running) equal the gradients of ONE process on the concatenated batch,
  * all ranks hold bit-identical reduced gradients,
  * after FusedAdamW steps the replicas are still bit-identical.
TEST INFRASTRUCTURE: no oracle involved -- the checker is the product's own single-rank path, which the parity tests
pin against the oracle.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from polyp_image_generator_b200 import DDPMScheduler, FusedAdamW, UNet2DModel
    from polyp_image_generator_b200.ddp import DistributedDataParallel
    from polyp_image_generator_b200.graphs import GraphedTrainStep
    from polyp_image_generator_b200.model import polyp_unet_config
    from polyp_image_generator_b200.training import mse_loss

    S, per = 64, 2
    B = per * world
    cfg = polyp_unet_config(S)
    cfg["block_out_channels"] = (128, 128, 128, 128, 256, 256)     # halo conv at 64 / pair kernel, attention at 4x4
    torch.manual_seed(0)
    model = UNet2DModel(**cfg).to(dev).train()
    single = UNet2DModel(**cfg).to(dev).train()
    g = torch.Generator().manual_seed(1)
    clean = (torch.rand(B, 3, S, S, generator=g) * 2 - 1).to(dev)
    noise = torch.randn(B, 3, S, S, generator=g).to(dev)
    t = torch.randint(0, 1000, (B,), generator=g).to(dev)
    sched = DDPMScheduler()
    ddp = DistributedDataParallel(model, bucket_cap_mb=4.0)          # several buckets: the overlap path is exercised
    single.load_state_dict(model.state_dict())
    sl = slice(rank * per, (rank + 1) * per)

    def mark(what):
        print(f"[rank {rank}] {what}", file=sys.stderr, flush=True)

    def flat_grads(m):
        return torch.cat([p.grad.reshape(-1).float() for p in m.parameters()])

    def rel(a, b):
        return ((a - b).norm() / (b.norm() + 1e-30)).item()

    mark("setup done")
    # ---- reference: one process, whole batch ----
    noisy = sched.add_noise(clean, noise, t)
    mse_loss(single(noisy, t, return_dict=False)[0], noise).backward()
    g_ref = flat_grads(single)

    mark("reference done")
    # ---- eager DDP ----
    loss = mse_loss(ddp(noisy[sl], t[sl], return_dict=False)[0], noise[sl])
    loss.backward()
    torch.cuda.synchronize()
    g_ddp = flat_grads(model)
    r_eager = rel(g_ddp, g_ref)
    gathered = [torch.empty_like(g_ddp) for _ in range(world)]
    dist.all_gather(gathered, g_ddp)
    same = all(torch.equal(gathered[0], x) for x in gathered)

    # ---- CUDA-graph DDP step (lr = 0: weights stay put, gradients of the replay are comparable) ----
    # (drop the eager autograd graph first: while it lives, the parameters' AccumulateGrad nodes stay bound to the stream
    # of that eager step -- the legacy default stream -- and a capture that reuses them is refused by CUDA)
    mark("eager DDP done")
    del loss
    opt = FusedAdamW(model.parameters(), lr=0.0, weight_decay=0.0, max_grad_norm=1.0)
    model.zero_grad(set_to_none=True)
    step = GraphedTrainStep(ddp, sched, opt, clean[sl].shape, max_grad_norm=1.0, warmup_iters=2,
                            warmup_batch=(clean[sl], noise[sl], t[sl]))
    mark("graph captured")
    loss_g = step(clean[sl], noise[sl], t[sl])
    torch.cuda.synchronize()
    mark("graph replayed")
    g_graph = flat_grads(model)
    r_graph = rel(g_graph, g_ref)
    # mean of the rank losses == whole-batch loss
    lt = loss_g.detach().clone().reshape(1)
    dist.all_reduce(lt, op=dist.ReduceOp.AVG)
    loss_ref = mse_loss(single(noisy, t, return_dict=False)[0].detach(), noise)

    mark("loss reduced")
    # ---- replicas stay identical through real optimizer steps ----
    opt2 = FusedAdamW(model.parameters(), lr=1e-3, max_grad_norm=1.0)
    for _ in range(3):
        model.zero_grad(set_to_none=True)
        mse_loss(ddp(noisy[sl], t[sl], return_dict=False)[0], noise[sl]).backward()
        opt2.step()
    torch.cuda.synchronize()
    mark("3 eager optimizer steps done")
    w = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    ws = [torch.empty_like(w) for _ in range(world)]
    dist.all_gather(ws, w)
    replicas_same = all(torch.equal(ws[0], x) for x in ws)

    ok = r_eager < 1e-2 and r_graph < 1e-2 and same and replicas_same and \
        abs(lt.item() - loss_ref.item()) < 2e-3 * abs(loss_ref.item())
    if rank == 0:
        print(f"nccl 2-rank vs 1-rank whole gradient: eager rel {r_eager:.3e}, graph rel {r_graph:.3e}, "
              f"ranks bit-identical {same}, replicas bit-identical after 3 steps {replicas_same}, "
              f"loss {lt.item():.6f} vs {loss_ref.item():.6f}", flush=True)
        print("NCCL_PARITY_OK" if ok else "NCCL_PARITY_FAIL", flush=True)
    okt = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)          # every rank leaves with the same exit code
    code = 0 if int(okt.item()) == 1 else 1
    mark(f"exit code {code}")
    sys.stdout.flush()
    sys.stderr.flush()
    # no destroy_process_group(): tearing the communicator down while a CUDA graph that captured its collectives is
    # still alive blocked both ranks until the test's timeout on this stack (torch 2.11 / NCCL 2.28.9)
    os._exit(code)


if __name__ == "__main__":
    main()
