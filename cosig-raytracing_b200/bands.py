"""Tile sharding of one frame over ranks: full-width bands of `band_rows` rows, band b belongs to rank
b % world; a rank's "local rows" are its bands packed in order.  The same arithmetic lives on the device
(csrc/rtb_device.cuh: band_local_rows / band_global_row); this module is the host side used by bench.py and the tests.

Two gathers to rank 0:
  * fused (default on NVLink): every rank's resolve kernel stores its pixels straight into rank 0's frame through a CUDA-IPC
    peer mapping (RayTracer.frame_export / frame_import) — no collective runs at all;
  * `gather_bands`: each rank renders its bands packed (RTB_OUT_COMPACT) and torch.distributed gathers them (NCCL on GPUs,
    gloo in the CPU tests); rank 0 scatters them into the frame.
"""
from __future__ import annotations

import numpy as np


def owned_rows(height: int, rank: int, world: int, band_rows: int = 32) -> np.ndarray:
    rows = np.arange(height)
    return rows[(rows // band_rows) % max(1, world) == rank] if world > 1 else rows


def local_row_count(height: int, rank: int, world: int, band_rows: int = 32) -> int:
    if world <= 1:
        return height
    n_bands = (height + band_rows - 1) // band_rows
    owned = (n_bands - rank + world - 1) // world if n_bands > rank else 0
    if owned == 0:
        return 0
    rows = owned * band_rows
    last_end = (rank + (owned - 1) * world + 1) * band_rows
    return rows - max(0, last_end - height)


def gather_bands(mine, height: int, width: int, rank: int, world: int, band_rows: int = 32, dst: int = 0):
    """mine: uint8 tensor [local_rows, width, 4] (this rank's packed bands, on the backend's device).  Returns the
    assembled [height, width, 4] frame on rank `dst`, None elsewhere."""
    import torch
    import torch.distributed as dist
    counts = [local_row_count(height, r, world, band_rows) for r in range(world)]
    assert mine.shape[0] == counts[rank]
    cap = max(counts)
    padded = torch.zeros((cap, width, 4), dtype=torch.uint8, device=mine.device)
    padded[:counts[rank]] = mine
    if rank == dst:
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.gather(padded, parts, dst=dst)
        frame = torch.empty((height, width, 4), dtype=torch.uint8, device=mine.device)
        for r in range(world):
            idx = torch.from_numpy(owned_rows(height, r, world, band_rows)).to(mine.device)
            frame[idx] = parts[r][:counts[r]]
        return frame
    dist.gather(padded, None, dst=dst)
    return None


def host_ring_copy_plan(height: int, width: int, rank: int, world: int, band_rows: int = 8) -> dict:
    """What a rank of a host-ring group (rtb_group_create_host) copies into the frame's slot, in bytes of the row-major RGBA8 frame:
    one strided copy — `full` pieces of `piece` bytes, the first at `first`, one every `pitch` — plus, when the rank's last band is
    cut short by the frame's end, one linear copy of `tail_bytes` at `tail_offset`.  Mirrors api.cu: group_host_begin (the C code is
    what runs; this restatement lets the CPU tests check that the ranks' plans tile a frame exactly once)."""
    row_bytes = width * 4
    need = row_bytes * height
    if world <= 1:
        return dict(first=0, pitch=need, piece=need, full=1, tail_offset=0, tail_bytes=0)
    band_bytes = row_bytes * band_rows
    n_bands = (height + band_rows - 1) // band_rows
    owned = (n_bands - rank + world - 1) // world if n_bands > rank else 0
    if owned == 0:
        return dict(first=0, pitch=0, piece=0, full=0, tail_offset=0, tail_bytes=0)
    last_band = rank + (owned - 1) * world
    last_short = (last_band + 1) * band_rows > height
    full = owned - 1 if last_short else owned
    off = last_band * band_bytes
    return dict(first=rank * band_bytes, pitch=world * band_bytes, piece=band_bytes, full=full,
                tail_offset=off if last_short else 0, tail_bytes=need - off if last_short else 0)


def apply_copy_plan(plan: dict, src: np.ndarray, dst: np.ndarray) -> int:
    """Executes a host_ring_copy_plan between two flat uint8 views of a frame; returns the bytes moved."""
    moved = 0
    for i in range(plan["full"]):
        a = plan["first"] + i * plan["pitch"]
        dst[a:a + plan["piece"]] = src[a:a + plan["piece"]]
        moved += plan["piece"]
    if plan["tail_bytes"]:
        a = plan["tail_offset"]
        dst[a:a + plan["tail_bytes"]] = src[a:a + plan["tail_bytes"]]
        moved += plan["tail_bytes"]
    return moved
