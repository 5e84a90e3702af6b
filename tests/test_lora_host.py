"""LoRA host logic on CPU (emulated op contracts): injection, key grammar, freezing, forward/backward wiring of the
adapter folded into the projection GEMMs, merge, save/load -- against the oracle's peft restatement."""
import os
import pytest
import torch

import oracle


def _cfg(S=32):
    cfg = oracle.polyp_unet_config(S)
    cfg["block_out_channels"] = (64, 64, 128, 128, 128, 128)
    return cfg


def _pair(dropout=0.0, targets=("to_q", "to_k", "to_v", "to_out.0"), r=8, alpha=8):
    from polyp_image_generator_b200 import LoraConfig, UNet2DModel
    torch.manual_seed(0)
    om = oracle.UNet2DModel(**_cfg())
    m = UNet2DModel(**_cfg())
    m.load_state_dict(om.state_dict())
    oracle.add_adapter(om, oracle.LoraConfig(r=r, lora_alpha=alpha, target_modules=list(targets),
                                             lora_dropout=dropout, init_lora_weights="gaussian"))
    m.add_adapter(LoraConfig(r=r, lora_alpha=alpha, target_modules=list(targets), lora_dropout=dropout,
                             init_lora_weights="gaussian"))
    # same adapter weights on both sides; make B non-zero so the branch matters
    sd = {k: torch.randn_like(v) * 0.05 for k, v in oracle.lora_state_dict(om).items()}
    assert om.load_state_dict(sd, strict=False).unexpected_keys == []
    assert m.load_state_dict(sd, strict=False).unexpected_keys == []
    return m, om


def test_injection_structure_and_keys(emu_backend):
    from polyp_image_generator_b200.lora import lora_state_dict, recover_lora_modules
    m, om = _pair()
    assert list(m.state_dict().keys()) == list(om.state_dict().keys())
    trainable = [n for n, p in m.named_parameters() if p.requires_grad]
    assert trainable and all("lora_" in n for n in trainable)
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == \
        sum(p.numel() for p in om.parameters() if p.requires_grad)
    sd = lora_state_dict(m)
    assert len(sd) == 48 and all(v.device.type == "cpu" for v in sd.values())
    assert recover_lora_modules(sd) == oracle.recover_lora_modules(oracle.lora_state_dict(om))
    assert "mid_block.attentions.0.to_out.0.lora_B.default.weight" in sd
    with pytest.raises(ValueError, match="already exists"):
        from polyp_image_generator_b200 import LoraConfig
        m.add_adapter(LoraConfig())


def test_unknown_targets_raise(emu_backend):
    from polyp_image_generator_b200 import LoraConfig, UNet2DModel
    m = UNet2DModel(**_cfg())
    with pytest.raises(ValueError, match="not found"):
        m.add_adapter(LoraConfig(target_modules=["proj_in"]))
    with pytest.raises(NotImplementedError):
        m.add_adapter(LoraConfig(target_modules=["linear_1"]))       # time-embedding MLP linears: not a supported target


@pytest.mark.parametrize("targets", [("to_q", "to_k", "to_v", "to_out.0"), ("to_q", "to_v")])
def test_lora_forward_backward_match_oracle(emu_backend, targets):
    m, om = _pair(targets=targets)
    m.train()
    om.train()
    x, t = torch.randn(2, 3, 32, 32), torch.tensor([3, 600])
    y, yo = m(x, t).sample, om(x, t).sample
    assert ((y - yo).norm() / yo.norm()).item() < 1e-5
    tgt = torch.randn_like(y)
    torch.nn.functional.mse_loss(y, tgt).backward()
    torch.nn.functional.mse_loss(yo, tgt).backward()
    og = dict(om.named_parameters())
    n_checked = 0
    # the 1-token mid-block attention has analytically zero q/k gradients: compare against the overall scale
    tot = sum(p.grad.norm() ** 2 for p in om.parameters() if p.grad is not None) ** 0.5
    for n, p in m.named_parameters():
        if not p.requires_grad:
            assert p.grad is None, n
            continue
        g = og[n].grad
        assert ((p.grad - g).norm() / (g.norm() + 1e-4 * tot)).item() < 2e-3, n
        n_checked += 1
    assert n_checked == 2 * len(targets) * 6


def test_merge_matches_oracle_and_unmerged_forward(emu_backend):
    from polyp_image_generator_b200.lora import merge_adapter, unmerge_adapter
    m, om = _pair()
    m.eval()
    om.eval()
    x, t = torch.randn(1, 3, 32, 32), torch.tensor([77])
    with torch.no_grad():
        y_unmerged = m(x, t).sample
    w_before = m.mid_block.attentions[0].to_q.base_layer.weight.detach().clone()
    merge_adapter(m)
    oracle.merge_adapter(om)
    for (n, p), (_, po) in zip(m.named_parameters(), om.named_parameters()):
        assert torch.allclose(p, po, rtol=0, atol=1e-5), n        # north_star: merged weights within 1e-5
    with torch.no_grad():
        y_merged = m(x, t).sample
    assert ((y_merged - y_unmerged).norm() / y_unmerged.norm()).item() < 1e-5
    unmerge_adapter(m)
    assert torch.allclose(m.mid_block.attentions[0].to_q.base_layer.weight, w_before, atol=1e-6)


def test_dropout_branch_is_consistent_between_forward_and_backward(emu_backend):
    """With lora_dropout > 0 the mask is regenerated (not stored): finite-difference check of dA through the model."""
    m, _ = _pair(dropout=0.3)
    m.train()
    x, t = torch.randn(1, 3, 32, 32), torch.tensor([500])
    A = m.up_blocks[1].attentions[2].to_v.lora_A["default"].weight     # 4-token attention close to the output
    from polyp_image_generator_b200.lora import GemmLora
    tgt = torch.randn(1, 3, 32, 32)

    def loss_at(delta):
        GemmLora._seed_counter = 100           # replay the same masks
        torch.manual_seed(9)
        with torch.no_grad():
            A.add_(delta)
            l = torch.nn.functional.mse_loss(m(x, t).sample, tgt).item()
            A.sub_(delta)
        return l

    GemmLora._seed_counter = 100
    torch.manual_seed(9)
    torch.nn.functional.mse_loss(m(x, t).sample, tgt).backward()
    d = torch.randn_like(A) * 0.2
    fd = (loss_at(d) - loss_at(-d)) / 2
    an = (A.grad * d).sum().item()
    assert abs(an) > 1e-6 and fd == pytest.approx(an, rel=0.15)
    # eval mode: no dropout
    m.eval()
    with torch.no_grad():
        y1, y2 = m(x, t).sample, m(x, t).sample
    assert torch.equal(y1, y2)


def test_accumulating_lora_epoch_matches_the_reference_loop(emu_backend):
    """train_with_lora_all_classes.py:123-176 restated (training.train_epoch_with_accumulation) against the same loop
    written over the oracle: loss / accumulation_steps, clip over ALL unet parameters, step + zero_grad + lr step every
    accumulation_steps batches, the trailing batch's gradients left unapplied."""
    from polyp_image_generator_b200 import DDPMScheduler, LoraConfig, UNet2DModel
    from polyp_image_generator_b200.training import get_cosine_schedule_with_warmup, train_epoch_with_accumulation
    cfg = oracle.polyp_unet_config(32)
    cfg["block_out_channels"] = (64, 64, 64, 64, 128, 128)
    tg = ["to_q", "to_k", "to_v", "to_out.0"]
    torch.manual_seed(0)
    om = oracle.UNet2DModel(**cfg)
    m = UNet2DModel(**cfg)
    m.load_state_dict(om.state_dict())
    oracle.add_adapter(om, oracle.LoraConfig(r=4, lora_alpha=4, target_modules=tg, init_lora_weights="gaussian"))
    m.add_adapter(LoraConfig(r=4, lora_alpha=4, target_modules=tg, init_lora_weights="gaussian"))
    sd = {k: torch.randn_like(v) * 0.05 for k, v in oracle.lora_state_dict(om).items()}
    om.load_state_dict(sd, strict=False)
    m.load_state_dict(sd, strict=False)
    opt = torch.optim.AdamW([p for p in m.parameters() if p.requires_grad], lr=1e-3)
    oopt = torch.optim.AdamW([p for p in om.parameters() if p.requires_grad], lr=1e-3)
    sch, osch = get_cosine_schedule_with_warmup(opt, 1, 4), get_cosine_schedule_with_warmup(oopt, 1, 4)
    g = torch.Generator().manual_seed(3)
    batches = [torch.randn(2, 3, 32, 32, generator=g).clamp(-1, 1) for _ in range(3)]
    draws = [(torch.randn(2, 3, 32, 32, generator=g), torch.randint(0, 1000, (2,), generator=g)) for _ in range(3)]
    it = iter(draws)
    avg = train_epoch_with_accumulation(m, DDPMScheduler(), opt, batches, lr_scheduler=sch, accumulation_steps=2,
                                        draw=lambda x: next(it))
    # the reference loop over the oracle objects
    ns, total = oracle.DDPMScheduler(), 0.0
    oopt.zero_grad()
    for step, (x, (noise, t)) in enumerate(zip(batches, draws)):
        loss = torch.nn.functional.mse_loss(om(ns.add_noise(x, noise, t), t).sample, noise) / 2
        loss.backward()
        if (step + 1) % 2 == 0:
            torch.nn.utils.clip_grad_norm_(list(om.parameters()), 1.0)
            oopt.step()
            oopt.zero_grad()
            osch.step()
        total += loss.item() * 2
    assert avg == pytest.approx(total / 3, rel=1e-5)
    assert sch.get_last_lr() == osch.get_last_lr()
    osd = oracle.lora_state_dict(om)
    from polyp_image_generator_b200.lora import lora_state_dict
    for k, v in lora_state_dict(m).items():
        assert torch.allclose(v, osd[k], rtol=1e-4, atol=2e-6), k
    # the third batch's gradients are still pending, as in the reference
    assert any(p.grad is not None and p.grad.abs().sum() > 0 for p in m.parameters() if p.requires_grad)


def test_lora_weights_file_round_trip(emu_backend, tmp_path):
    """save_lora_weights / load_lora_weights (train_with_lora_all_classes.py:29-38) and the module recovery of
    get_lorarized_layers.py on the file they write."""
    from polyp_image_generator_b200 import LoraConfig, UNet2DModel
    from polyp_image_generator_b200.lora import load_lora_weights, recover_lora_modules, save_lora_weights
    cfg = oracle.polyp_unet_config(32)
    cfg["block_out_channels"] = (64, 64, 64, 64, 128, 128)
    tg = ["to_q", "to_k", "to_v", "to_out.0"]
    torch.manual_seed(1)
    a, b = UNet2DModel(**cfg), UNet2DModel(**cfg)
    b.load_state_dict(a.state_dict())
    for m in (a, b):
        m.add_adapter(LoraConfig(r=8, lora_alpha=8, target_modules=tg, lora_dropout=0.3, init_lora_weights="gaussian"))
    with torch.no_grad():
        for n, p in a.named_parameters():
            if "lora_" in n:
                p.normal_(0, 0.1)
    path = save_lora_weights(a, str(tmp_path / "lora_AD"))
    assert os.path.basename(path) == "lora_weights.pth"
    sd = torch.load(path)
    assert len(sd) == 48 and all("lora_" in k and v.device.type == "cpu" for k, v in sd.items())
    mods = recover_lora_modules(sd)
    assert len(mods) == 24 and "mid_block.attentions.0.to_out.0" in mods and "down_blocks.4.attentions.1.to_q" in mods
    load_lora_weights("cpu", b, str(tmp_path / "lora_AD"))
    x, t = torch.randn(1, 3, 32, 32), torch.tensor([7])
    a.eval(); b.eval()
    with torch.no_grad():
        assert torch.equal(a(x, t).sample, b(x, t).sample)


def test_polyp_generator_model_wrapper_lora_plus_unfreezing(emu_backend, capsys):
    """PolypGeneratorModel.py:13-64: construction, add_lora_config's report, unfreeze_layers by name substring -- and
    the backward program honouring the resulting mix of frozen / adapter / unfrozen parameters (the
    'unconditional_with_lora_and_unfreezing' runs of the reference)."""
    from polyp_image_generator_b200 import LoraConfig, PolypGeneratorModel
    from polyp_image_generator_b200.training import mse_loss
    torch.manual_seed(21)
    wrap = PolypGeneratorModel("cpu", pretrained=False, add_lora=True, image_size=32)
    m = wrap.get_model()
    assert sum(p.numel() for p in m.parameters()) == 113_673_219 and m.config.sample_size == 32
    with pytest.raises(NotImplementedError):
        PolypGeneratorModel("cpu", pretrained=True, add_lora=False)
    om = oracle.UNet2DModel(**oracle.polyp_unet_config(32))
    om.load_state_dict(m.state_dict())
    tg = ["to_q", "to_k", "to_v", "to_out.0"]
    wrap.add_lora_config(LoraConfig(r=8, lora_alpha=8, target_modules=tg, init_lora_weights="gaussian"))
    assert "Trainable params of unet: 196608 / 113869827 (0.17%)" in capsys.readouterr().out
    oracle.add_adapter(om, oracle.LoraConfig(r=8, lora_alpha=8, target_modules=tg, init_lora_weights="gaussian"))
    sd = {k: torch.randn_like(v) * 0.05 for k, v in oracle.lora_state_dict(om).items()}
    om.load_state_dict(sd, strict=False)
    m.load_state_dict(sd, strict=False)
    unfreeze = ["conv_out", "up_blocks.5.resnets.2.conv2", "conv_norm_out"]
    wrap.unfreeze_layers(unfreeze)
    for n, p in om.named_parameters():
        if any(x in n for x in unfreeze):
            p.requires_grad = True
    m.eval(); om.eval()
    x, t, tgt = torch.randn(1, 3, 32, 32), torch.tensor([321]), torch.randn(1, 3, 32, 32)
    mse_loss(m(x, t).sample, tgt).backward()
    torch.nn.functional.mse_loss(om(x, t).sample, tgt).backward()
    og = dict(om.named_parameters())
    tot = sum(p.grad.norm() ** 2 for p in om.parameters() if p.grad is not None) ** 0.5
    seen = 0
    for n, p in m.named_parameters():
        if p.requires_grad:
            g = og[n].grad      # (q / k adapters of the 1-token mid-block attention have exactly zero gradient)
            assert p.grad is not None and ((p.grad - g).norm() / (g.norm() + 1e-6 * tot)).item() < 2e-3, n
            seen += 1
        else:
            assert p.grad is None, n
    assert seen == 48 + 6


@pytest.mark.parametrize("targets", [("time_emb_proj",), ("to_q", "to_k", "to_v", "to_out.0", "time_emb_proj")])
def test_lora_on_time_emb_proj_matches_oracle(emu_backend, targets):
    """config_diffusion.py:37 lists time_emb_proj among the candidate LoRA targets: the adapters on the 32 per-block
    time-embedding projections (fp32 side computation on the [batch, 512] embedding) against the oracle's peft
    restatement -- keys, forward, every adapter gradient, merge, and the fp32-faithful inference program."""
    from polyp_image_generator_b200.lora import merge_adapter, unmerge_adapter
    m, om = _pair(targets=targets)
    assert list(m.state_dict().keys()) == list(om.state_dict().keys())
    n_res = sum(1 for n, _ in om.named_modules() if n.endswith("time_emb_proj"))
    m.train()
    om.train()
    x, t = torch.randn(2, 3, 32, 32), torch.tensor([3, 600])
    y, yo = m(x, t).sample, om(x, t).sample
    assert ((y - yo).norm() / yo.norm()).item() < 1e-5
    tgt = torch.randn_like(y)
    torch.nn.functional.mse_loss(y, tgt).backward()
    torch.nn.functional.mse_loss(yo, tgt).backward()
    og = dict(om.named_parameters())
    tot = sum(p.grad.norm() ** 2 for p in om.parameters() if p.grad is not None) ** 0.5
    n_temb = 0
    for n, p in m.named_parameters():
        if not p.requires_grad:
            assert p.grad is None, n
            continue
        g = og[n].grad
        assert ((p.grad - g).norm() / (g.norm() + 1e-4 * tot)).item() < 2e-3, n
        n_temb += "time_emb_proj" in n
    assert n_temb == 2 * n_res and n_res >= 20
    # fp32-faithful inference program carries the adapters too
    m.eval()
    om.eval()
    m.inference_precision = "fp32"
    with torch.no_grad():
        assert ((m(x, t).sample - om(x, t).sample).norm() / yo.norm()).item() < 1e-4
    m.inference_precision = "bf16"
    # merge: weights within 1e-5 of the oracle's, forward unchanged, unmerge restores
    with torch.no_grad():
        y_un = m(x, t).sample
    merge_adapter(m)
    oracle.merge_adapter(om)
    for (n, p), (_, po) in zip(m.named_parameters(), om.named_parameters()):
        assert torch.allclose(p, po, rtol=0, atol=1e-5), n
    with torch.no_grad():
        assert ((m(x, t).sample - y_un).norm() / y_un.norm()).item() < 1e-5
    unmerge_adapter(m)
    with torch.no_grad():
        assert ((m(x, t).sample - y_un).norm() / y_un.norm()).item() < 1e-5


def test_lora_on_time_emb_proj_dropout_trains_and_is_off_in_eval(emu_backend):
    m, _ = _pair(dropout=0.3, targets=("time_emb_proj",))
    x, t = torch.randn(2, 3, 32, 32), torch.tensor([3, 600])
    m.eval()
    with torch.no_grad():
        a, b = m(x, t).sample, m(x, t).sample
    assert torch.equal(a, b)                       # no dropout outside training
    m.train()
    torch.manual_seed(1)
    y1 = m(x, t).sample
    y2 = m(x, t).sample
    assert not torch.equal(y1, y2)                 # fresh masks per forward
    y2.square().mean().backward()
    gs = [p.grad for n, p in m.named_parameters() if "lora_" in n]
    assert gs and all(g is not None and torch.isfinite(g).all() for g in gs)


def test_lora_on_time_emb_proj_with_a_trainable_time_embedding_mlp(emu_backend):
    """PolypGeneratorModel.unfreeze_layers (PolypGeneratorModel.py:61-64) can make the time-embedding MLP trainable next
    to the adapters: the adapters' path contributes to d_emb (they read SiLU(emb) as well)."""
    m, om = _pair(targets=("time_emb_proj", "to_q"))
    for mod in (m, om):
        for n, p in mod.named_parameters():
            if n.startswith("time_embedding."):
                p.requires_grad_(True)
        mod.train()
    x, t = torch.randn(2, 3, 32, 32), torch.tensor([3, 600])
    y, yo = m(x, t).sample, om(x, t).sample
    tgt = torch.randn_like(y)
    torch.nn.functional.mse_loss(y, tgt).backward()
    torch.nn.functional.mse_loss(yo, tgt).backward()
    og = dict(om.named_parameters())
    tot = sum(p.grad.norm() ** 2 for p in om.parameters() if p.grad is not None) ** 0.5
    seen = 0
    for n, p in m.named_parameters():
        if p.requires_grad:
            g = og[n].grad
            assert ((p.grad - g).norm() / (g.norm() + 1e-4 * tot)).item() < 2e-3, n
            seen += n.startswith("time_embedding.")
    assert seen == 4
