"""Cross-check the LoRA backward against (a) the oracle and (b) our own full-weight wgrad on the merged model."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn.functional as F
import oracle
from polyp_image_generator_b200 import LoraConfig, UNet2DModel
from polyp_image_generator_b200.lora import merge_adapter
from polyp_image_generator_b200.training import mse_loss

dev = "cuda"
for S, B in ((32, 3), (64, 4)):
    cfg = dict(oracle.polyp_unet_config(S))
    cfg["block_out_channels"] = (64, 64, 128, 128, 256, 256)
    torch.manual_seed(7)
    om = oracle.UNet2DModel(**cfg)
    m = UNet2DModel(**cfg)
    m.load_state_dict(om.state_dict())
    m2 = UNet2DModel(**cfg)
    m2.load_state_dict(om.state_dict())
    tg = ["to_q", "to_k", "to_v", "to_out.0"]
    oracle.add_adapter(om, oracle.LoraConfig(r=8, lora_alpha=8, target_modules=tg, init_lora_weights="gaussian"))
    m.add_adapter(LoraConfig(r=8, lora_alpha=8, target_modules=tg, init_lora_weights="gaussian"))
    sd = {k: torch.randn_like(v) * 0.05 for k, v in oracle.lora_state_dict(om).items()}
    om.load_state_dict(sd, strict=False)
    m.load_state_dict(sd, strict=False)
    m.to(dev).train()
    x, t, nz = torch.randn(B, 3, S, S), torch.randint(0, 1000, (B,)), torch.randn(B, 3, S, S)
    pred = m(x.to(dev), t.to(dev)).sample
    mse_loss(pred, nz.to(dev)).backward()
    pred_o = om(x, t).sample
    F.mse_loss(pred_o, nz).backward()
    # merged full model on our kernels: dW' -> expected LoRA grads
    oracle.merge_adapter(om)
    base_sd = {k.replace(".base_layer", ""): v for k, v in om.state_dict().items() if "lora_" not in k}
    m2.load_state_dict(base_sd)
    m2.to(dev).train()
    pred2 = m2(x.to(dev), t.to(dev)).sample
    mse_loss(pred2, nz.to(dev)).backward()
    g2 = {n: p.grad for n, p in m2.named_parameters()}
    og = {n: p.grad for n, p in om.named_parameters()}
    print(f"S={S}: pred rel lora-vs-oracle {((pred.cpu()-pred_o).norm()/pred_o.norm()).item():.2e}  merged-vs-oracle {((pred2.cpu()-pred_o).norm()/pred_o.norm()).item():.2e}")
    rows = []
    for n, p in m.named_parameters():
        if not p.requires_grad:
            continue
        g_or = og[n]
        base = n.split(".lora_")[0]
        dW = g2[base + ".weight"].float().cpu()
        A = sd[base + ".lora_A.default.weight"]; Bm = sd[base + ".lora_B.default.weight"]
        exp = (Bm.t() @ dW) if ".lora_A." in n else (dW @ A.t())      # scaling = 1
        def r(a, b): return ((a - b).norm() / (b.norm() + 1e-30)).item()
        rows.append((n, g_or.norm().item(), r(p.grad.cpu(), g_or), r(exp, g_or), r(p.grad.cpu(), exp)))
    rows.sort(key=lambda z: -z[2])
    for n, gn, e1, e2, e3 in rows[:10]:
        print(f"  {n:62s} |g|={gn:.2e} lora-vs-oracle {e1:.3f}  mergedW-vs-oracle {e2:.3f}  lora-vs-mergedW {e3:.3f}")
