"""Batched parameter sweep = the reference's SampleParameters loop (heteroclinic.cpp:429-491) over many GPUs.

One process per GPU.  Fronts are independent, so the sweep is sharded statically (front_id mod world,
SURVEY.md 8e) with NO collective during compute; the only exchange is one all-gather of the fixed-size
result records at the end (NCCL over NVLink on GPUs, gloo in the CPU tests of this plumbing).
torch is used for device buffers, streams and torch.distributed only.
"""
import ctypes as C

import numpy as np
import torch

from . import abi


def desc_array_to_tensor(descs, device=None, pin=False):
    """Copy a ctypes FrontDesc array into a uint8 tensor (optionally pinned / on device)."""
    raw = np.frombuffer(descs, dtype=np.uint8).copy()
    t = torch.from_numpy(raw)
    if pin:
        t = t.pin_memory()
    if device is not None:
        t = t.to(device, non_blocking=False)
    return t


def tensor_to_results(t):
    """uint8 tensor (host) -> ctypes FrontResult array."""
    raw = t.cpu().numpy().tobytes()
    m = len(raw) // C.sizeof(abi.FrontResult)
    return (abi.FrontResult * m).from_buffer_copy(raw)


def local_indices(n_total, rank, world):
    return list(range(rank, n_total, world))


def padded_shard_len(n_total, world):
    return (n_total + world - 1) // world


def gather_records(local, n_total, rank, world, record_bytes, group=None):
    """All-gather the per-rank result records (uint8 tensor [m_local*record_bytes], on the device the backend
    wants) and undo the interleaved sharding: returns [n_total*record_bytes] in front_id order on every rank."""
    import torch.distributed as dist
    m_pad = padded_shard_len(n_total, world)
    buf = torch.zeros(m_pad * record_bytes, dtype=torch.uint8, device=local.device)
    buf[: local.numel()] = local
    if world == 1:
        out = buf.view(1, m_pad, record_bytes)
    else:
        out = torch.empty(world * m_pad * record_bytes, dtype=torch.uint8, device=local.device)
        dist.all_gather_into_tensor(out, buf, group=group)          # the single collective of the path
        out = out.view(world, m_pad, record_bytes)
    # front f lives at rank f % world, slot f // world
    full = out.permute(1, 0, 2).reshape(m_pad * world, record_bytes)[:n_total]
    return full.reshape(-1)


def counts_from_records(rec_bytes_tensor):
    """(status, zero_count, failure_count) int32 columns from gathered result records."""
    rb = C.sizeof(abi.FrontResult)
    a = rec_bytes_tensor.cpu().numpy().reshape(-1, rb)[:, :12].copy().view(np.int32)
    return a[:, 0], a[:, 1], a[:, 2]


def write_plot_parameters(path, params, counts):
    """plot_parameters.txt as SampleParameters writes it (heteroclinic.cpp:474-486): `c12 c23 count`, precision 8,
    negative codes collapsed to -1 (heteroclinic.cpp:469-470)."""
    with open(path, "w") as f:
        for p, c in zip(params, counts):
            f.write("%.8g %.8g %d\n" % (p[3], p[4], c if c >= 0 else -1))
