"""PIE-Bench-shaped sweep (BASELINE.json configs[4]): N synthetic 512x512 images x {P2P, MasaCtrl, PnP, Pix2Pix-zero}, DDIM inversion,
sharded image-wise over the GPUs of one box — the serial loops of p2p/test.py:114-181, masactrl/test.py, pnp/test.py and
pix2pix-zero/test.py with image i on rank i mod G, one process per GPU, no collective on the hot path (runner.run_sharded gathers the
small per-image records at the end).

    python tools/sweep.py --images 700                                          one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep.py --images 700

What is synthetic: the images (seeded uniform uint8, like bench.py), the prompt pairs (drawn from templates so that, as in PIE-Bench,
some pairs have equal word counts -> AttentionReplace and others do not -> AttentionRefine, p2p/test.py:120-123) and the nine category
ids of p2p/test.py:114. "Direct inversion", which BASELINE.json's config text mentions, does not exist in the reference (SURVEY.md
fact 0.4): every method runs the reference's DDIM inversion. Objects are kept across images (controller.retarget() / editor.reset()),
so with --graphs the UNet forwards replay captured CUDA graphs after the first image of each kind.
Each record carries a CRC of the result images: a G-GPU sweep must reproduce the 1-GPU sweep bit for bit (tests/test_gpu_e2e.py).
P2P, MasaCtrl and PnP are bit-reproducible as they are (every kernel of this library is deterministic, forward-only torch ops are);
Pix2Pix-zero differentiates through the UNet, and cuDNN's default backward-data convolutions are not — pass --deterministic
(torch.backends.cudnn.deterministic) when its CRCs must match too.
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402
import torch         # noqa: E402

CATEGORIES = [0, 1, 2, 3, 4, 6, 7, 8, 9]          # p2p/test.py:114
SUBJECTS = ["cat", "dog", "horse", "squirrel", "bird", "fox", "hippopotamus", "rabbit"]
PLACES = ["on a bench", "in the garden", "near the river", "on a table", "by the lake"]
STYLES = ["a photo of", "a painting of", "a watercolor drawing of", "a sketch of"]
METHODS = ("p2p", "masactrl", "pnp", "pix2pix_zero")


def item(i: int, side: int = 512):
    """Image i of the synthetic dataset: (category, uint8 image, source prompt, target prompt)."""
    rng = np.random.default_rng(9000 + i)
    image = rng.integers(0, 256, size=(side, side, 3), dtype=np.uint8)
    a, b = rng.choice(len(SUBJECTS), 2, replace=False)
    place, style = PLACES[rng.integers(len(PLACES))], STYLES[rng.integers(len(STYLES))]
    source = f"{style} a {SUBJECTS[a]} {place}"
    if i % 3 == 2:      # a third of the pairs add words: different word counts -> refine
        target = f"{style} a large {SUBJECTS[a]} sleeping {place}"
    else:               # word swap: equal word counts -> replace
        target = f"{style} a {SUBJECTS[b]} {place}"
    return CATEGORIES[i % len(CATEGORIES)], image, source, target


class Worker:
    """One rank's pipeline replica and the kept per-method objects."""

    def __init__(self, device, ddim_steps: int, graphs: bool, config: str = "sd15", dtype=torch.bfloat16, deterministic: bool = False):
        if deterministic:
            torch.backends.cudnn.deterministic = True
        import image_editing_framework_b200 as pkg
        from image_editing_framework_b200 import p2p, masactrl, pnp, pix2pix_zero  # noqa: F401  (bind the sub-packages on pkg)
        from image_editing_framework_b200.ddim import ddim_inversion
        from image_editing_framework_b200.standin import make_pipeline
        from image_editing_framework_b200.standin.unet import sd15_config, tiny_config, AttnProcessor
        self.pkg, self.dev, self.steps, self.graphs, self.dtype = pkg, device, ddim_steps, graphs, dtype
        AttnProcessor.use_sdpa = True       # layers no method hooks run what diffusers runs on a GPU
        cfg = sd15_config() if config == "sd15" else tiny_config()
        with torch.device(device):
            self.pipe = make_pipeline(cfg, seed=0, device=device, dtype=dtype)
        if dtype != torch.float32:
            self.pipe.unet.to(memory_format=torch.channels_last)
        self.side = cfg.sample_size * 8
        self.inv = ddim_inversion()
        self.inv.graphs = graphs
        self.plain = pkg.masactrl.AttentionBase()
        self.kept = {}

    def _sync(self):
        if torch.device(self.dev).type == "cuda":
            torch.cuda.synchronize(self.dev)

    def _keep(self, key, make):
        if key not in self.kept:
            self.kept[key] = make()
        return self.kept[key]

    # ---- shared first stage: VAE encode + DDIM inversion with the source prompt (every */test.py does this) ---------------------
    def invert(self, image, source):
        pkg, pipe = self.pkg, self.pipe
        pipe.scheduler.set_timesteps(self.steps)
        latent = self.inv.image2latent(pipe, image, self.dev, self.dtype)
        self.plain.reset()
        pkg.masactrl.regiter_attention_editor_diffusers(pipe, self.plain)     # routes the inversion's attention through the kernels
        try:
            trajectory, _ = self.inv.ddim_inversion_loop(pipe, latent, [source])
        finally:
            pkg.masactrl.unregister_attention_control(pipe, self.plain)
        return trajectory

    def p2p(self, x_t, source, target):
        pkg, pipe = self.pkg, self.pipe
        prompts = [source, target]
        replace = len(source.split(" ")) == len(target.split(" "))                  # p2p/test.py:120-123
        cls = pkg.p2p.AttentionReplace if replace else pkg.p2p.AttentionRefine
        ctrl = self.kept.get(("p2p_ctrl", replace))
        if ctrl is None:
            ctrl = self.kept[("p2p_ctrl", replace)] = cls(prompts, pipe.tokenizer, self.steps, 0.8, 0.6, device=self.dev)
        else:
            ctrl.retarget(prompts, pipe.tokenizer)
        editor = self._keep(("p2p", replace), lambda: pkg.p2p.P2P(pipe, self.steps, graphs=self.graphs))
        editor.height = editor.width = self.side      # 512 for the SD-1.5 stand-in (the class constant), smaller for the debug config
        try:
            images, _ = editor.text2image_ldm_stable(pipe, prompts, ctrl, latent=x_t, num_inference_steps=self.steps, guidance_scale=7.5)
        finally:
            pkg.p2p.unregister_attention_control(pipe, ctrl)
        return images

    def masactrl(self, x_t, source, target):
        pkg, pipe = self.pkg, self.pipe
        editor = self._keep("masa_editor", lambda: pkg.masactrl.MutualSelfAttentionControl(4 if self.steps > 8 else 1, 10, total_steps=self.steps))
        editor.reset()
        pkg.masactrl.regiter_attention_editor_diffusers(pipe, editor)
        try:
            sampler = self._keep("masa", lambda: pkg.masactrl.MasaCtrl(pipe, self.steps, graphs=self.graphs))
            images, _ = sampler([source, target], latents=torch.cat([x_t, x_t]), guidance_scale=7.5, num_inference_steps=self.steps,
                                height=self.side, width=self.side)
        finally:
            pkg.masactrl.unregister_attention_control(pipe, editor)
        return images

    def pnp(self, x_t, source, target):
        sampler = self._keep("pnp", lambda: self.pkg.pnp.PnP(self.pipe, self.steps, graphs=self.graphs))
        return sampler([source, target], num_inference_steps=self.steps, guidance_scale=7.5, latents=x_t, pnp_attn_t=0.5, pnp_f_t=0.8,
                       height=self.side, width=self.side)

    def pix2pix_zero(self, x_t, source, target):
        pkg, pipe = self.pkg, self.pipe
        editor = self._keep("p2z", lambda: pkg.pix2pix_zero.P2P_Zero(pipe, self.steps, graphs=self.graphs))
        try:
            return editor([source, target], num_inference_steps=self.steps, guidance_scale=7.5, latents=x_t.clone(), guidance_amount=0.1,
                          height=self.side, width=self.side)[1]
        finally:
            pkg.pix2pix_zero.restore_original_processors(pipe.unet, editor.original_processors)

    def run(self, i: int, methods):
        category, image, source, target = item(i, self.side)
        rec = {"category": category, "source": source, "target": target, "crc": {}, "ms": {}}
        with contextlib.redirect_stdout(io.StringIO()), torch.no_grad():
            self._sync()
            t0 = time.perf_counter()
            x_t = self.invert(image, source)[-1]
            self._sync()
            rec["ms"]["inversion"] = (time.perf_counter() - t0) * 1e3
        for m in methods:
            with contextlib.redirect_stdout(io.StringIO()):
                t0 = time.perf_counter()
                ctx = torch.enable_grad() if m == "pix2pix_zero" else torch.no_grad()   # its guidance pass differentiates through the UNet
                with ctx:
                    images = getattr(self, m)(x_t, source, target)
                self._sync()
                rec["ms"][m] = (time.perf_counter() - t0) * 1e3
            images = np.ascontiguousarray(np.asarray(images))
            assert images.dtype == np.uint8 and images.shape[-3:] == (self.side, self.side, 3), (images.dtype, images.shape)
            rec["crc"][m] = zlib.crc32(images.tobytes())
        return rec


def sweep(worker: Worker, n_images: int, methods, rank: int = None, world: int = None):
    """This rank's share of the sweep -> {image index: record}. rank / world default to the process group's."""
    from image_editing_framework_b200 import runner
    if rank is None:
        return runner.run_sharded(lambda i: worker.run(i, methods), n_images)
    return {i: worker.run(i, methods) for i in runner.shard_indices(n_images, rank, world)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=700)
    ap.add_argument("--methods", default=",".join(METHODS))
    ap.add_argument("--ddim-steps", type=int, default=50)
    ap.add_argument("--config", default="sd15", choices=["sd15", "tiny"])
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--deterministic", action="store_true", help="cuDNN deterministic algorithms (needed for bit-equal Pix2Pix-zero results)")
    ap.add_argument("--out", default=None, help="per-image records (JSON lines), written by rank 0")
    args = ap.parse_args()
    import torch.distributed as dist
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    methods = [m for m in args.methods.split(",") if m]
    worker = Worker(dev, args.ddim_steps, not args.no_graphs, args.config, deterministic=args.deterministic)
    # untimed: the first image of each kind pays cuDNN autotuning and the graph captures
    warm = {i: worker.run(i, methods) for i in (0, 2)}
    del warm
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    records = sweep(worker, args.images, methods)
    torch.cuda.synchronize()
    wall = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(wall, op=dist.ReduceOp.MAX)
    if rank == 0:
        wall_s = wall.item()
        per_method = {m: round(sum(r["ms"][m] for r in records.values()) / len(records), 1) for m in ["inversion"] + methods}
        line = {"sweep": f"{args.images} synthetic {worker.side}x{worker.side} images x {methods}, DDIM inversion ({args.ddim_steps} steps) + edit "
                         f"({args.ddim_steps} steps), {args.config} stand-in, bf16", "n_gpus": world, "images": len(records), "wall_s": round(wall_s, 2),
                "images_per_s": round(len(records) / wall_s, 4), "edits_per_s": round(len(records) * len(methods) / wall_s, 4),
                "mean_ms_per_image_on_its_gpu": per_method, "cuda_graphs": not args.no_graphs,
                "replace_pairs": sum(1 for r in records.values() if len(r["source"].split(" ")) == len(r["target"].split(" "))),
                "crc_of_crcs": zlib.crc32(json.dumps([records[i]["crc"] for i in sorted(records)], sort_keys=True).encode())}
        print(json.dumps(line), flush=True)
        if args.out:
            with open(args.out, "w") as f:
                for i in sorted(records):
                    f.write(json.dumps({"image": i, **records[i]}) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
