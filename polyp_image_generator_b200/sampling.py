"""Per-class sampling bookkeeping around `DDPMPipeline`.

Call sites restated (paths relative to /root/reference/generator_model/):
  train_from_scratch.py:39-66          evaluate(config, epoch, pipeline, cls, imgs_to_generate)
  train_with_lora_per_class.py:59-88   evaluate(config, pipeline, cls, prompt, num_images, out_dir)
  train_with_lora_per_class.py:252-290 "top-up": count the files already in samples/<cls>, generate the difference

Contract kept from the reference: batches of `config.eval_batch_size` (last one ragged), batch b is drawn from
`torch.Generator('cpu').manual_seed(config.seed + b)`, files are `<n>.png` with n 1-based in generation order, written to
`<output_dir>/samples/<cls>/`.  What changes is only where the time goes:
  * with world > 1 rank r generates batches r, r+world, ... (ddp.shard_sampling_batches): same seeds, same file names,
    no communication -- the union over ranks is exactly the single-process image set;
  * the pipeline returns device uint8 NHWC (the [0,1]*255 epilogue is a kernel) and PNG encoding + disk writes run on a
    small thread pool, so the next batch's 1000 UNet forwards never wait for file I/O.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional

import torch

from .ddp import shard_sampling_batches


def _save_png(arr, path: str) -> str:
    from PIL import Image
    Image.fromarray(arr.squeeze() if arr.shape[-1] == 1 else arr).save(path)
    return path


def generate_images(config, pipeline, out_dir: str, num_images: int, *, first_index: int = 1, rank: int = 0,
                    world: int = 1, num_inference_steps: int = 1000, io_threads: int = 4,
                    verbose: bool = True) -> List[str]:
    """The `while total < num_images` loop of both reference `evaluate` functions.  Returns the paths this rank wrote
    (in generation order).  `first_index` is the number of the first file (the reference always starts at 1)."""
    os.makedirs(out_dir, exist_ok=True)
    bs = int(config.eval_batch_size)
    if bs <= 0:
        raise ValueError("config.eval_batch_size must be positive")
    written: List[str] = []
    futures = []
    with ThreadPoolExecutor(max_workers=max(1, io_threads)) as pool:
        for batch_id, start, count in shard_sampling_batches(int(num_images), bs, rank, world):
            gen = torch.Generator(device="cpu").manual_seed(int(config.seed) + batch_id)
            u8 = pipeline(batch_size=count, generator=gen, num_inference_steps=num_inference_steps,
                          output_type="uint8").images                    # device uint8 NHWC
            host = u8.to("cpu", non_blocking=False).numpy()
            for i in range(count):
                path = os.path.join(out_dir, f"{first_index + start + i}.png")
                futures.append(pool.submit(_save_png, host[i], path))
                written.append(path)
            if verbose:
                print(f"   Saved {start + count} images")
        for f in futures:
            f.result()
    return written


def evaluate(config, epoch, pipeline, cls: str, imgs_to_generate: int, *, rank: int = 0, world: int = 1,
             num_inference_steps: int = 1000, verbose: bool = True) -> List[str]:
    """train_from_scratch.py:39-66 (mlflow logging of ten sample files is the caller's business)."""
    cls_dir = os.path.join(config.output_dir, "samples", cls)
    paths = generate_images(config, pipeline, cls_dir, imgs_to_generate, rank=rank, world=world,
                            num_inference_steps=num_inference_steps, verbose=verbose)
    if verbose:
        print(f"  {imgs_to_generate} images saved at {cls_dir}")
    return paths


def count_samples(samples_dir: str) -> int:
    """train_with_lora_per_class.py:266-267: number of regular files already in samples/<cls>."""
    if not os.path.isdir(samples_dir):
        return 0
    return sum(1 for f in os.scandir(samples_dir) if f.is_file())


def top_up(config, pipeline, cls: str, target: int, folder: Optional[str] = None, *, continue_numbering: bool = False,
           rank: int = 0, world: int = 1, num_inference_steps: int = 1000, verbose: bool = True) -> List[str]:
    """train_with_lora_per_class.py:262-290: if samples/<cls> holds fewer than `target` files, generate the difference
    (all of them when the directory does not exist).  The reference restarts file numbering AND batch seeds at 1 / 0
    for the top-up run, so it re-creates (overwrites) 1.png ... <missing>.png; that behaviour is kept by default.
    `continue_numbering=True` appends after the existing files instead (seeds still start at config.seed)."""
    folder = folder if folder is not None else config.output_dir
    out_dir = os.path.join(folder, "samples", cls)
    have = count_samples(out_dir)
    if have >= target:
        return []
    missing = target - have
    first = have + 1 if continue_numbering else 1
    return generate_images(config, pipeline, out_dir, missing, first_index=first, rank=rank, world=world,
                           num_inference_steps=num_inference_steps, verbose=verbose)


def per_class_resume_plan(folder: str, classes, targets):
    """The resume / skip decision of train_with_lora_per_class.py:252-293 for each class, as data:

        "train"     no `lora_<cls>` + `model_<cls>` pair in `folder` yet: train (and sample) from scratch     (:292-)
        "generate"  trained, but `samples/<cls>` does not exist: generate all `target` images              (:281-290)
        "top_up"    trained, `samples/<cls>` holds fewer files than `target`: generate the difference     (:265-278)
        "done"      trained and enough samples

    Returns [(cls, action, n_images_to_generate)] in the order of `classes`."""
    present = set(os.listdir(folder)) if os.path.isdir(folder) else set()
    plan = []
    for cls, target in zip(classes, targets):
        if f"lora_{cls}" in present and f"model_{cls}" in present:
            sdir = os.path.join(folder, "samples", cls)
            if os.path.exists(sdir):
                have = count_samples(sdir)
                plan.append((cls, "top_up", target - have) if have < target else (cls, "done", 0))
            else:
                plan.append((cls, "generate", target))
        else:
            plan.append((cls, "train", target))
    return plan
