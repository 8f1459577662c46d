"""Stand-ins for the rest of a diffusers Stable-Diffusion pipeline: DDIM scheduler, CLIP tokenizer /
text encoder, VAE and the pipeline object the reference's drivers receive.

NOT product code (test / bench infrastructure, see standin/unet.py). Semantics restated from memory of
diffusers 0.27 and cross-checked against the reference's call sites:
  scheduler config   p2p/edit_real.py:58-69   (scaled_linear 0.00085..0.012, 1000 steps, steps_offset=1,
                                               set_alpha_to_one=False, clip_sample=False)
  scheduler.step     p2p/model/sd_utils.py:76; alphas_cumprod / final_alpha_cumprod / timesteps: inversion/ddim.py:11-13,28
  tokenizer surface  p2p/model/seq_aligner.py:108-109,139; p2p/model/ptp_utils.py:42; p2p/model/sd_utils.py:42-53
The real CLIP vocabulary is not available offline; `WordPieceTokenizer` is a deterministic
stand-in that yields multi-token words, so CLIP's actual merges are never exercised.
"""
from __future__ import annotations

import contextlib
from types import SimpleNamespace
from typing import List, Optional, Sequence, Union

import numpy as np
import torch
import torch.nn as nn

from .unet import UNet2DConditionModel, UNetConfig, _Output, tiny_config


class DDIMScheduler:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False, set_alpha_to_one=False,
                 steps_offset=1, num_train_timesteps=1000):
        assert beta_schedule == "scaled_linear" and not clip_sample
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, steps_offset=steps_offset, beta_start=beta_start,
                                      beta_end=beta_end, set_alpha_to_one=set_alpha_to_one, clip_sample=clip_sample)
        self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64) + self.config.steps_offset
        self.timesteps = torch.from_numpy(ts).to(device) if device is not None else torch.from_numpy(ts)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step(self, model_output, timestep, sample, eta: float = 0.0, return_dict: bool = True, **kw):
        assert eta == 0.0
        prev = timestep - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[timestep]
        a_prev = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        direction = (1 - a_prev) ** 0.5 * model_output
        return _Output(prev_sample=a_prev ** 0.5 * x0 + direction, pred_original_sample=x0)


class WordPieceTokenizer:
    """Deterministic stand-in for CLIPTokenizer: words are cut into pieces of at most `piece` characters and every
    piece maps to an invertible integer id, so `decode([id])` returns the piece text."""
    _ALPHABET = "abcdefghijklmnopqrstuvwxyz0123456789'-.,!?_"

    def __init__(self, model_max_length: int = 77, piece: int = 4):
        self.model_max_length = model_max_length
        self.piece = piece
        self.bos_token_id, self.eos_token_id = 49406, 49407
        self._base = len(self._ALPHABET) + 1

    def _piece_id(self, s: str) -> int:
        v = 0
        for ch in s:
            i = self._ALPHABET.find(ch.lower())
            v = v * self._base + (i + 1 if i >= 0 else self._base - 1)
        return v  # < 45**4 ~ 4.1M, never collides with bos/eos for 4-char pieces over this alphabet? keep them apart:

    def encode(self, text: str) -> List[int]:
        ids = [self.bos_token_id]
        for word in text.split(" "):
            if not word:
                continue
            for i in range(0, len(word), self.piece):
                ids.append(100000 + self._piece_id(word[i:i + self.piece]))
        ids.append(self.eos_token_id)
        return ids

    def decode(self, ids: Sequence[int]) -> str:
        out = []
        for t in ids:
            t = int(t)
            if t in (self.bos_token_id, self.eos_token_id):
                out.append("<|startoftext|>" if t == self.bos_token_id else "<|endoftext|>")
                continue
            v, s = t - 100000, ""
            while v > 0:
                v, r = divmod(v, self._base)
                s = (self._ALPHABET[r - 1] if r - 1 < len(self._ALPHABET) else "#") + s
            out.append(s)
        return " ".join(out)

    def __call__(self, text: Union[str, List[str]], padding="max_length", max_length: Optional[int] = None, truncation=True,
                 return_tensors="pt"):
        texts = [text] if isinstance(text, str) else list(text)
        L = max_length or self.model_max_length
        rows = []
        for t in texts:
            ids = self.encode(t)[:L]
            if len(ids) == L:
                ids[-1] = self.eos_token_id
            rows.append(ids + [self.eos_token_id] * (L - len(ids)))
        return SimpleNamespace(input_ids=torch.tensor(rows, dtype=torch.int64))


class TextEncoder(nn.Module):
    """Hash-embedding + one linear layer; returns a tuple like CLIPTextModel (`[0]` = last hidden state)."""

    def __init__(self, dim: int = 768, max_len: int = 77, buckets: int = 4096):
        super().__init__()
        self.buckets = buckets
        self.tok = nn.Embedding(buckets, dim)
        self.pos = nn.Embedding(max_len, dim)
        self.mix = nn.Linear(dim, dim)

    def forward(self, input_ids):
        x = self.tok(input_ids % self.buckets) + self.pos(torch.arange(input_ids.shape[1], device=input_ids.device))[None]
        return (self.mix(torch.tanh(x)),)


class VAE(nn.Module):
    def __init__(self, latent_channels: int = 4):
        super().__init__()
        self.config = SimpleNamespace(scaling_factor=0.18215)
        self.enc = nn.Conv2d(3, latent_channels, 8, stride=8)
        self.dec = nn.ConvTranspose2d(latent_channels, 3, 8, stride=8)

    def encode(self, image):
        return {"latent_dist": SimpleNamespace(mean=self.enc(image))}

    def decode(self, latents):
        return {"sample": self.dec(latents)}


class StableDiffusionPipeline:
    """Carries the attributes the reference drivers read: unet, vae, tokenizer, text_encoder, scheduler, device."""

    def __init__(self, unet: UNet2DConditionModel, scheduler: Optional[DDIMScheduler] = None, text_dim: Optional[int] = None):
        self.unet = unet
        self.scheduler = scheduler or DDIMScheduler()
        self.tokenizer = WordPieceTokenizer()
        self.text_encoder = TextEncoder(text_dim or unet.config.cross_attention_dim)
        self.vae = VAE(unet.config.in_channels)
        self.vae_scale_factor = 8

    def to(self, device=None, dtype=None):
        for m in (self.unet, self.text_encoder, self.vae):
            m.to(device=device, dtype=dtype)
        return self

    @property
    def device(self):
        return self.unet.device

    @property
    def _execution_device(self):
        return self.unet.device

    def encode_prompt(self, prompt, device, num_images_per_prompt=1, do_classifier_free_guidance=True, negative_prompt=None,
                      prompt_embeds=None, negative_prompt_embeds=None, lora_scale=None, **kw):
        prompts = [prompt] if isinstance(prompt, str) else list(prompt)
        ids = self.tokenizer(prompts, padding="max_length", max_length=self.tokenizer.model_max_length, truncation=True).input_ids
        pe = self.text_encoder(ids.to(device))[0]
        ne = None
        if do_classifier_free_guidance:
            neg = [negative_prompt or ""] * len(prompts)
            nid = self.tokenizer(neg, padding="max_length", max_length=self.tokenizer.model_max_length).input_ids
            ne = self.text_encoder(nid.to(device))[0]
        return pe, ne

    def prepare_latents(self, batch_size, num_channels, height, width, dtype, device, generator, latents=None):
        if latents is None:
            latents = torch.randn((batch_size, num_channels, height // 8, width // 8), generator=generator, dtype=dtype).to(device)
        return latents.to(device) * self.scheduler.init_noise_sigma

    def prepare_extra_step_kwargs(self, generator, eta):
        return {}

    @contextlib.contextmanager
    def progress_bar(self, total=None):
        yield SimpleNamespace(update=lambda *a, **k: None)


def make_pipeline(config: Optional[UNetConfig] = None, seed: int = 0, device="cpu", dtype=torch.float32) -> StableDiffusionPipeline:
    """Random-init pipeline (no weights are available offline). Deterministic in `seed`."""
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        unet = UNet2DConditionModel(config or tiny_config())
        pipe = StableDiffusionPipeline(unet)
        # break the symmetry of default init a little so attention maps are not near-uniform
        with torch.no_grad():
            for name, p in unet.named_parameters():
                if name.endswith(("to_q.weight", "to_k.weight")):
                    p.mul_(3.0)
    finally:
        torch.random.set_rng_state(g)
    for m in (pipe.unet, pipe.text_encoder, pipe.vae):
        m.eval().requires_grad_(False)
    return pipe.to(device=device, dtype=dtype)
