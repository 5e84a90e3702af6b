"""Multi-process coverage of the data-parallel path on CPU: world size 2, `gloo`, the CPU emulation of the
op contracts standing in for the kernels.  What is checked is the HOST logic of ddp.py -- parameter broadcast, bucketed
all-reduce of the flat gradient arena driven by the backward program's progress callbacks, the LoRA extra-gradient path,
identical parameters after optimizer steps -- and that the rank shards of the sampling bookkeeping reproduce the
single-process image set."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup(rank, world, port):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from emu_ops import EmuOps
    from polyp_image_generator_b200 import ops as ops_mod
    ops_mod.set_backend(EmuOps())


def _small_cfg():
    import oracle
    cfg = oracle.polyp_unet_config(32)
    cfg["block_out_channels"] = (64, 64, 64, 64, 128, 128)
    return cfg


def _flat_grads(model):
    return torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.requires_grad and p.grad is not None])


def _worker_full_training(rank, world, port):
    _setup(rank, world, port)
    try:
        from polyp_image_generator_b200 import DDPMScheduler, FusedAdamW, UNet2DModel
        from polyp_image_generator_b200.ddp import DistributedDataParallel
        from polyp_image_generator_b200.training import mse_loss, train_step
        torch.manual_seed(100 + rank)                       # ranks start with DIFFERENT weights ...
        model = UNet2DModel(**_small_cfg())
        net = DistributedDataParallel(model, bucket_cap_mb=0.25)   # ... small buckets: several progress-driven all-reduces
        ref = UNet2DModel(**_small_cfg())
        ref.load_state_dict(model.state_dict())
        w0 = [torch.empty_like(model.conv_in.weight) for _ in range(world)]
        dist.all_gather(w0, model.conv_in.weight.detach().contiguous())
        assert torch.equal(w0[0], w0[1]), "broadcast_parameters did not equalise the ranks"

        g = torch.Generator().manual_seed(7 + rank)          # each rank its own shard of the batch
        x = torch.randn(2, 3, 32, 32, generator=g)
        t = torch.randint(0, 1000, (2,), generator=g)
        tgt = torch.randn(2, 3, 32, 32, generator=g)
        mse_loss(net(x, t, return_dict=False)[0], tgt).backward()
        mse_loss(ref(x, t, return_dict=False)[0], tgt).backward()      # local, un-reduced gradient of the same shard
        local = _flat_grads(ref)
        both = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(both, local)
        want = (both[0] + both[1]) / world
        got = _flat_grads(model)
        assert got.shape == want.shape
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-7), float((got - want).abs().max())
        # every parameter received a gradient, and it is a view into the flat arena that was all-reduced
        assert all(p.grad is not None for p in model.parameters())

        # two optimisation steps through the reference loop body keep the replicas bit-identical
        opt = FusedAdamW(net.parameters(), lr=1e-3, max_grad_norm=1.0)
        opt.zero_grad()
        sched = DDPMScheduler()
        for _ in range(2):
            noise = torch.randn(2, 3, 32, 32, generator=g)
            train_step(net, sched, opt, x, noise, t)
        flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        others = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(others, flat)
        assert torch.equal(others[0], others[1]), "replicas diverged after optimizer steps"
        assert not torch.equal(flat, torch.cat([p.detach().reshape(-1) for p in ref.parameters()]))
    finally:
        dist.destroy_process_group()


def _worker_lora(rank, world, port):
    _setup(rank, world, port)
    try:
        from polyp_image_generator_b200 import LoraConfig, UNet2DModel
        from polyp_image_generator_b200.ddp import DistributedDataParallel
        from polyp_image_generator_b200.training import mse_loss
        torch.manual_seed(5)
        model = UNet2DModel(**_small_cfg())
        model.add_adapter(LoraConfig(r=4, lora_alpha=4, target_modules=["to_q", "to_k", "to_v", "to_out.0"],
                                     init_lora_weights="gaussian"))
        with torch.no_grad():                                # B = 0 at init would make every A-gradient vanish
            for n, p in model.named_parameters():
                if "lora_B" in n:
                    p.normal_(0, 0.05, generator=torch.Generator().manual_seed(11))
        ref = UNet2DModel(**_small_cfg())
        ref.add_adapter(LoraConfig(r=4, lora_alpha=4, target_modules=["to_q", "to_k", "to_v", "to_out.0"],
                                   init_lora_weights="gaussian"))
        ref.load_state_dict(model.state_dict())
        net = DistributedDataParallel(model)
        g = torch.Generator().manual_seed(70 + rank)
        x, t, tgt = torch.randn(2, 3, 32, 32, generator=g), torch.randint(0, 1000, (2,), generator=g), \
            torch.randn(2, 3, 32, 32, generator=g)
        model.eval()
        ref.eval()                                           # adapter dropout off: deterministic comparison
        for m in (model, ref):
            for p in m.parameters():
                p.grad = None
        mse_loss(net(x, t, return_dict=False)[0], tgt).backward()
        mse_loss(ref(x, t, return_dict=False)[0], tgt).backward()
        names = [n for n, p in model.named_parameters() if p.requires_grad]
        assert names and all("lora_" in n for n in names)
        local = _flat_grads(ref)
        both = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(both, local)
        want = (both[0] + both[1]) / world
        got = _flat_grads(model)
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-7), float((got - want).abs().max())
        assert all(p.grad is None for n, p in model.named_parameters() if not p.requires_grad)
    finally:
        dist.destroy_process_group()


def _worker_sampling(rank, world, port, out_dir):
    _setup(rank, world, port)
    try:
        from types import SimpleNamespace
        from polyp_image_generator_b200 import DDPMPipeline, DDPMScheduler, UNet2DModel
        from polyp_image_generator_b200.sampling import evaluate
        torch.manual_seed(3)
        cfg = _small_cfg()
        cfg["block_out_channels"] = (64, 64, 64, 64, 64, 64)
        pipe = DDPMPipeline(unet=UNet2DModel(**cfg), scheduler=DDPMScheduler())
        conf = SimpleNamespace(output_dir=out_dir, eval_batch_size=2, seed=0)
        evaluate(conf, 0, pipe, "AD", 5, rank=rank, world=world, num_inference_steps=2, verbose=False)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_ddp_gradients_are_the_mean_of_the_rank_gradients_and_replicas_stay_identical():
    mp.spawn(_worker_full_training, args=(2, _free_port()), nprocs=2, join=True)


@pytest.mark.timeout(600)
def test_ddp_lora_gradients_all_reduced():
    mp.spawn(_worker_lora, args=(2, _free_port()), nprocs=2, join=True)


@pytest.mark.timeout(600)
def test_sharded_sampling_reproduces_the_single_process_images(tmp_path, emu_backend):
    import numpy as np
    from PIL import Image
    from types import SimpleNamespace
    from polyp_image_generator_b200 import DDPMPipeline, DDPMScheduler, UNet2DModel
    from polyp_image_generator_b200.sampling import evaluate
    mp.spawn(_worker_sampling, args=(2, _free_port(), str(tmp_path / "two")), nprocs=2, join=True)
    torch.manual_seed(3)
    cfg = _small_cfg()
    cfg["block_out_channels"] = (64, 64, 64, 64, 64, 64)
    pipe = DDPMPipeline(unet=UNet2DModel(**cfg), scheduler=DDPMScheduler())
    conf = SimpleNamespace(output_dir=str(tmp_path / "one"), eval_batch_size=2, seed=0)
    paths = evaluate(conf, 0, pipe, "AD", 5, num_inference_steps=2, verbose=False)
    assert len(paths) == 5
    for p in paths:
        q = os.path.join(str(tmp_path / "two"), "samples", "AD", os.path.basename(p))
        assert os.path.exists(q), q
        assert np.array_equal(np.asarray(Image.open(p)), np.asarray(Image.open(q)))
