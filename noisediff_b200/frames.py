"""Full-frame noise synthesis: tile-grid producer + asynchronous ``.npy`` writer (BASELINE config 4).

What the reference does for one long-exposure frame (``dataloader/dataset.py:203-281`` + ``Trainer.test``,
``models/trainer_diffusion.py:240-325``): cut the packed 4x1424x2128 frame into overlapping ``ps x ps`` crops (88 at
ps = 256, 24 at ps = 512), attach each crop's position map and the ISO/ratio index, run ``diffusion.sample`` batch by batch
and ``np.save`` every generated crop as ``<clean>+<noisy|clean>+<x>_<y>.npy`` (float32, (4, ps, ps)) under
``<save_folder>/npy/generated``.  The crops are independent outputs — the reference does no stitching.

Here the crops and position maps are built on the device from the resident frame (no DataLoader round trip), ranks take
contiguous slices of the crop list (``tiles.shard``; no collective), and the files are written by a background thread from
pinned host buffers so that disk I/O never stalls the next batch's chain.
"""
from __future__ import annotations

import contextlib
import os
import queue
import threading
from typing import Dict, List, NamedTuple, Optional, Sequence, Tuple

import numpy as np
import torch

from . import tiles


@contextlib.contextmanager
def rank_noise_stream(diffusion, rank: int, world_size: int):
    """Independent noise per rank, derived INSIDE the library.  Ranks of a sharded run are usually seeded identically (DDP
    practice), which would give every rank the same x_T / z_t stream for its own crops — correlated synthetic noise across
    shards.  One draw from the caller's CPU generator (identical on identically seeded ranks; it also makes successive calls
    differ) is combined with the rank by ``tiles.rank_seed``, and the sampling inside the context runs on forked CPU + CUDA
    generators seeded with the result — both ``noise_source='torch'`` (x_T, z_t from the CUDA generator) and ``'philox'`` (base
    seed from the CPU generator) become rank-distinct, and the caller's generators continue as if only that one draw happened.
    A single-rank run is left untouched (bit-identical to calling ``diffusion.sample`` directly)."""
    if world_size <= 1:
        yield
        return
    base = int(torch.randint(0, 2 ** 40, (1,)).item())
    dev = diffusion.device
    devices = [dev] if dev.type == "cuda" else []
    with torch.random.fork_rng(devices=devices):
        torch.manual_seed(tiles.rank_seed(base, rank))       # seeds the CPU and every CUDA generator of this process
        yield


def crop_batch(clean_frame: torch.Tensor, origins: Sequence[Tuple[int, int]], ps: int, iso_ratio_idx: int) -> Dict[str, torch.Tensor]:
    """Condition dict for the crops at `origins` ((x, y) pairs) of one packed frame (4, H, W): what
    ``NoiseImageGenerationDataset.__getitem__`` yields per item (``dataset.py:242-281``), batched, on the frame's device."""
    if clean_frame.dim() != 3 or clean_frame.shape[0] != 4:
        raise ValueError(f"expected a packed (4, H, W) frame, got {tuple(clean_frame.shape)}")
    _, fh, fw = clean_frame.shape
    dev = clean_frame.device
    clean = torch.stack([clean_frame[:, y:y + ps, x:x + ps] for x, y in origins]).float().contiguous()
    if clean.shape[-2:] != (ps, ps):
        raise ValueError("crop origin outside the frame")
    pos = torch.stack([tiles.position_map(ps, ps, x, y, fh, fw, device=dev) for x, y in origins])
    idx = torch.full((len(origins),), int(iso_ratio_idx), dtype=torch.long, device=dev)
    return {"clean_img": clean, "position": pos, "iso_ratio_idx": idx}


def npy_name(clean_name: str, x: int, y: int, noisy_name: Optional[str] = None) -> str:
    """``Trainer.test`` file name (``models/trainer_diffusion.py:305-312``): both names lose their '.ARW' suffix."""
    c = clean_name.split(".ARW")[0]
    s = (noisy_name if noisy_name is not None else clean_name).split(".ARW")[0]
    return f"{c}+{s}+{int(x)}_{int(y)}.npy"


def dark_npy_name(index: int, iso: int, ratio: int, x: int, y: int) -> str:
    """``Trainer.test`` file name in ``--dark_frame`` mode (``models/trainer_diffusion.py:318-322``): a running item counter
    over the whole run instead of the frame names, ``%05d_<iso>_<ratio>+<x>_<y>.npy``."""
    return f"{int(index):05d}_{int(iso)}_{int(ratio)}+{int(x)}_{int(y)}.npy"


def parse_npy_name(name: str) -> Tuple[str, str, int, int]:
    """Inverse of :func:`npy_name`, the way the downstream consumer reads it (``dataloader/dataset_denoising.py:56-59,136-138``):
    ``clean+noisy+x_y.npy`` -> (clean, noisy, x, y)."""
    clean, noisy, coord = os.path.basename(name).split(".npy")[0].split("+")
    x, y = coord.split("_")
    return clean, noisy, int(x), int(y)


def consumer_subfolder(iso: int, ratio: int) -> str:
    """Folder the denoiser's synthetic-pair dataset expects per camera setting (``dataset_denoising.py:47-52``)."""
    return f"ISO{int(iso)}_Ratio{int(ratio)}"


def compose_noisy(clean_crop: torch.Tensor, noise: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """What the consumer does with a generated crop (``dataset_denoising.py:140-153``, without dark shading):
    noise clipped to [-1, 1], added to the clean crop, both clipped to [0, 1].  Returns (clean, noisy)."""
    if noise.is_cuda:      # on the GPU: one fused elementwise kernel of the library (ndiff_compose_noisy) on the generated batch
        import ctypes as C
        from . import _lib
        z = noise.detach().float().contiguous()
        c = clean_crop.detach().to(z.device).float().expand_as(z).contiguous()
        if z.numel() % 4 or (z.data_ptr() | c.data_ptr()) % 16:
            raise ValueError("compose_noisy: CUDA tensors must hold a multiple of 4 elements in 16-byte aligned storage")
        noisy, clean = torch.empty_like(z), torch.empty_like(c)
        with torch.cuda.device(z.device):
            _lib.check(_lib.lib().ndiff_compose_noisy(C.c_void_p(z.data_ptr()), C.c_void_p(c.data_ptr()), C.c_void_p(noisy.data_ptr()),
                                                      C.c_void_p(clean.data_ptr()), z.numel(),
                                                      C.c_void_p(torch.cuda.current_stream(z.device).cuda_stream)))
        return clean, noisy
    # host tensors: the consumer's own numpy lines, restated (plumbing on the caller's CPU data, no device involved)
    noisy = (noise.clamp(-1.0, 1.0).float() + clean_crop.float()).clamp(0.0, 1.0)
    return clean_crop.float().clamp(0.0, 1.0), noisy


class NpyWriter:
    """Background ``np.save`` of generated crops.  ``submit`` copies a finished batch to pinned host memory with a
    non-blocking D2H copy and returns; the worker waits for the copy's event and writes the files."""

    def __init__(self, folder: str, depth: int = 4):
        self.folder = folder
        os.makedirs(folder, exist_ok=True)
        self._q: "queue.Queue" = queue.Queue(maxsize=depth)
        self._err: Optional[BaseException] = None
        self.paths: List[str] = []
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def _run(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            host, event, names = item
            try:
                if event is not None:
                    event.synchronize()
                arr = host.numpy()
                for i, n in enumerate(names):
                    path = os.path.join(self.folder, n)
                    if os.path.dirname(n):                   # names may carry a sub-folder (one per camera setting)
                        os.makedirs(os.path.dirname(path), exist_ok=True)
                    np.save(path, arr[i])
                    self.paths.append(path)
            except BaseException as e:          # surfaced by close()
                self._err = e

    def submit(self, batch: torch.Tensor, names: Sequence[str]):
        if len(names) != batch.shape[0]:
            raise ValueError("one file name per item")
        if batch.is_cuda:
            host = torch.empty(batch.shape, dtype=torch.float32, pin_memory=True)
            host.copy_(batch.detach().float(), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(batch.device))
        else:
            host, ev = batch.detach().float().contiguous().clone(), None
        self._q.put((host, ev, list(names)))

    def close(self) -> List[str]:
        self._q.put(None)
        self._t.join()
        if self._err is not None:
            raise self._err
        return self.paths


@torch.inference_mode()
def synthesize_frame(diffusion, clean_frame: torch.Tensor, *, iso_ratio_idx: int, clean_name: str, save_folder: str,
                     noisy_name: Optional[str] = None, batch_size: int = 64, rank: int = 0, world_size: int = 1,
                     dark_frame: bool = False, iso: Optional[int] = None, ratio: Optional[int] = None,
                     index_base: int = 0) -> List[str]:
    """Generates this rank's share of the noise crops of one frame and writes them the way ``Trainer.test`` does.
    `diffusion` is a ``GaussianDiffusion`` (its ``image_size`` is the crop size).  Returns the written paths.
    ``dark_frame`` (ref :287-290,318-322): zero clean image, files named by a running counter (``index_base`` + the crop's position
    in the frame's grid, so ranks never collide) and the ISO / ratio, which must then be given."""
    ps = int(diffusion.image_size)
    dev = diffusion.device
    frame = clean_frame.to(dev)
    _, fh, fw = frame.shape
    origins = tiles.tile_origins(ps, fh, fw)
    my_range = tiles.shard(len(origins), world_size, rank)
    mine = [origins[i] for i in my_range]
    if dark_frame and (iso is None or ratio is None):
        raise ValueError("dark_frame file names carry the ISO and the ratio (trainer_diffusion.py:320-321): pass iso= and ratio=")
    # Trainer.test() layout by default; with (iso, ratio) the files go straight into the folder the denoiser's dataset globs
    folder = os.path.join(save_folder, consumer_subfolder(iso, ratio)) if iso is not None and ratio is not None \
        else os.path.join(save_folder, "npy", "generated")
    writer = NpyWriter(folder)
    try:
        with rank_noise_stream(diffusion, rank, world_size):
            for lo in range(0, len(mine), batch_size):
                part = mine[lo:lo + batch_size]
                cond = crop_batch(frame, part, ps, iso_ratio_idx)
                if dark_frame:                                   # ref :287-290: a zero clean image
                    cond["clean_img"] = torch.zeros_like(cond["clean_img"])
                out = diffusion.sample(batch_size=len(part), condition=cond)
                if dark_frame:
                    names = [dark_npy_name(index_base + my_range[lo + i], iso, ratio, x, y) for i, (x, y) in enumerate(part)]
                else:
                    names = [npy_name(clean_name, x, y, noisy_name) for x, y in part]
                writer.submit(out, names)
    finally:
        paths = writer.close()
    return paths


class FrameJob(NamedTuple):
    """One long-exposure frame to synthesise noise for: what one ``NoiseImageGenerationDataset`` file contributes
    (``dataloader/dataset.py:222-281``) plus where ``Trainer.test`` puts its crops (``models/trainer_diffusion.py:296-317``)."""
    clean_frame: torch.Tensor                 # packed (4, H, W), host (ideally pinned) or device
    iso_ratio_idx: int
    clean_name: str
    noisy_name: Optional[str] = None
    dark_frame: bool = False                  # ref :287-290: a zero clean image
    iso: Optional[int] = None                 # with (iso, ratio): files go to the consumer's ISO{iso}_Ratio{ratio}/ folder
    ratio: Optional[int] = None


def plan_crops(jobs: Sequence[FrameJob], ps: int) -> List[Tuple[int, int, int]]:
    """(job index, x, y) of every crop of every frame: frames in the given order, crops in the reference's grid order."""
    out = []
    for j, job in enumerate(jobs):
        _, fh, fw = job.clean_frame.shape
        out.extend((j, x, y) for x, y in tiles.tile_origins(ps, fh, fw))
    return out


def _job_folder(job: FrameJob) -> str:
    return consumer_subfolder(job.iso, job.ratio) if job.iso is not None and job.ratio is not None else os.path.join("npy", "generated")


@torch.inference_mode()
def synthesize_frames(diffusion, jobs: Sequence[FrameJob], *, save_folder: str, batch_size: int = 64, rank: int = 0,
                      world_size: int = 1) -> List[str]:
    """Several frames at once: the crops of ALL frames form one list, ranks take contiguous slices of it and sample it in FULL
    batches that may span frames.  One frame gives a rank of an 8-GPU box only 11 crops (88 / 8) — a batch size at which the 3x3
    convolutions at the 32x32 / 64x64 levels run far below their rate — whereas a list of frames keeps every
    engine at `batch_size` crops until the very last batch.  Crops stay independent outputs, written exactly as
    :func:`synthesize_frame` writes them; no collective."""
    ps = int(diffusion.image_size)
    dev = diffusion.device
    plan = plan_crops(jobs, ps)
    my_range = tiles.shard(len(plan), world_size, rank)
    mine = [plan[i] for i in my_range]
    if any(j.dark_frame and (j.iso is None or j.ratio is None) for j in jobs):
        raise ValueError("dark_frame file names carry the ISO and the ratio (trainer_diffusion.py:320-321): set FrameJob.iso / .ratio")
    writer = NpyWriter(save_folder)
    resident: Dict[int, torch.Tensor] = {}       # frames of the current batch on the device
    try:
        with rank_noise_stream(diffusion, rank, world_size):
            for lo in range(0, len(mine), batch_size):
                part = mine[lo:lo + batch_size]
                for j in [j for j in resident if j < part[0][0]]:
                    del resident[j]                   # the slice is ordered by frame: earlier frames are finished
                conds, names = [], []
                k = 0
                while k < len(part):                  # runs of crops from the same frame
                    j = part[k][0]
                    n = next((i for i, c in enumerate(part[k:]) if c[0] != j), len(part) - k)
                    run = [(x, y) for _, x, y in part[k:k + n]]
                    job = jobs[j]
                    if j not in resident:
                        resident[j] = job.clean_frame.to(dev, non_blocking=True)
                    cond = crop_batch(resident[j], run, ps, job.iso_ratio_idx)
                    if job.dark_frame:
                        cond["clean_img"] = torch.zeros_like(cond["clean_img"])
                    conds.append(cond)
                    if job.dark_frame:                # running counter over the whole job list = the crop's index in the plan
                        names += [os.path.join(_job_folder(job), dark_npy_name(my_range[lo + k + i], job.iso, job.ratio, x, y))
                                  for i, (x, y) in enumerate(run)]
                    else:
                        names += [os.path.join(_job_folder(job), npy_name(job.clean_name, x, y, job.noisy_name)) for x, y in run]
                    k += len(run)
                cond = {key: torch.cat([c[key] for c in conds]) for key in conds[0]}
                out = diffusion.sample(batch_size=len(part), condition=cond)
                writer.submit(out, names)
    finally:
        paths = writer.close()
    return paths
