"""Batched parameter sweep = the reference's SampleParameters loop (heteroclinic.cpp:429-491) over many GPUs.

One process per GPU.  Fronts are independent, so the sweep is sharded statically (front_id mod world,
SURVEY.md 8e) with NO collective during compute; the only exchange is one all-gather of the fixed-size
result records at the end (NCCL over NVLink on GPUs, gloo in the CPU tests of this plumbing).
torch is used for device buffers, streams and torch.distributed only.
"""
import ctypes as C
import os

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


def gather_records(local, n_total, rank, world, record_bytes, group=None, out=None, ordered=True):
    """All-gather the per-rank result records (uint8 tensor [m_local*record_bytes], on the device the backend wants): THE single
    collective of the path.  Front f lives at rank f % world, slot f // world (interleaved sharding).

    ordered=True : returns [n_total*record_bytes] in front order on every rank (one permuting copy after the collective).
    ordered=False: returns the collective's own layout, a [world, m_pad, record_bytes] view of `out` -- no copy at all; read front f
                   at [f % world, f // world] (`front_order` does it on the host).  `out` (world*m_pad*record_bytes bytes, same
                   device) can be preallocated and reused across sweeps; ranks whose shard is shorter than m_pad leave the
                   padding slot untouched."""
    import torch.distributed as dist
    m_pad = padded_shard_len(n_total, world)
    if world == 1:
        if local.numel() == m_pad * record_bytes:
            res = local.view(1, m_pad, record_bytes)
        else:
            res = torch.zeros(m_pad * record_bytes, dtype=torch.uint8, device=local.device); res[: local.numel()] = local
            res = res.view(1, m_pad, record_bytes)
    else:
        if local.numel() != m_pad * record_bytes:               # ragged last slot: pad this rank's contribution
            buf = torch.zeros(m_pad * record_bytes, dtype=torch.uint8, device=local.device)
            buf[: local.numel()] = local
            local = buf
        if out is None:
            out = torch.empty(world * m_pad * record_bytes, dtype=torch.uint8, device=local.device)
        dist.all_gather_into_tensor(out, local, group=group)
        res = out.view(world, m_pad, record_bytes)
    if not ordered:
        return res
    return res.permute(1, 0, 2).reshape(m_pad * world, record_bytes)[:n_total].reshape(-1)


def front_order(gathered, n_total):
    """[world, m_pad, record_bytes] (gather_records(..., ordered=False)) -> host uint8 array [n_total, record_bytes] in front order."""
    g = gathered.cpu().numpy()
    world, m_pad, rb = g.shape
    return np.ascontiguousarray(g.transpose(1, 0, 2).reshape(world * m_pad, rb)[:n_total])


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


# ----------------------------------------------------------------------------- SampleParameters (heteroclinic.cpp:429-491)

def _journal_key(descs, m):
    """Identity of a sweep for the crash journal: number of fronts + SHA-256 of the descriptor bytes (parameters, boxes, L_-,
    order / step / grid all live in the descriptors)."""
    import hashlib
    return "ccp-journal v2 m=%d sha256=%s" % (m, hashlib.sha256(bytes(np.frombuffer(descs, dtype=np.uint8))).hexdigest() if m else "empty")


def _read_journal(path, key):
    """Journal = one header line `# <key>` and lines `index code`.  A journal written for OTHER descriptors (different list, boxes,
    L_-, grid ...) or by an older version has another header and is discarded as a whole; a torn last line (crash while writing)
    is dropped."""
    done = {}
    try:
        with open(path) as f:
            text = f.read()
    except FileNotFoundError:
        return done
    lines = text.split("\n")
    if text and not text.endswith("\n"):
        lines = lines[:-1]
    if not lines or lines[0] != "# " + key:
        return done
    for ln in lines[1:]:
        parts = ln.split()
        if len(parts) == 2:
            done[int(parts[0])] = int(parts[1])
    return done


def SampleParameters(frame_count_batch, descs, params, path=None, chunk=1024, resume=True, code_of=None, unvalidated=False):
    """The sweep loop of SampleParameters (heteroclinic.cpp:447-471) over prepared descriptors, in chunks of `chunk`
    fronts per launch: Unstable_EigenValues[i] = code_of(result record), every negative code collapsed to -1
    (heteroclinic.cpp:469-470), then `plot_parameters.txt` in the reference's format (:474-486).

    In the reference the code is the return value of computeFrontAndConjugatePoints: -1 / -2 / -4 / -5 when the manifold conditions,
    the Krawczyk proof of the connecting orbit, the L_+ condition or the test below L_- fail, else frameDet's count.  The GPU hot path
    only delivers the frameDet part, so the caller must say which it wants:
      * `code_of(res) -> int` that folds in the outcome of the host stages (i), (ii), (iv) for that front (the C++ driver
        computeFrontAndConjugatePointsBatch of host/ccp_driver.hpp runs them all; its codes can be passed through here), or
      * `unvalidated=True`: the code is `frame_det_code` alone -- the frame count on inputs whose boxes were ASSUMED (e.g. the float
        shooting of ccp_setup_sweep).  The file is then named `plot_parameters_frame_count_only.txt` unless `path` is given: it has
        the reference's format but is not the reference's validated result.
    Neither given: ValueError.

    The reference keeps all counts in memory until the very end and loses the sweep on a crash; here every finished
    chunk is appended to `<path>.journal` (fsync'ed), a rerun with resume=True skips what the journal holds, and the
    journal is removed once the final file is written.  The journal starts with a header that identifies the sweep (number of fronts +
    SHA-256 of the descriptor bytes): a journal left by a different sweep is discarded, never merged.  `frame_count_batch` is
    `Engine.frame_count_batch` (the GPU call; there is no CPU fallback -- a failing call propagates).  Returns the list of counts.
    """
    from . import frame_det_code
    if code_of is None:
        if not unvalidated:
            raise ValueError("SampleParameters: pass code_of (a code that includes the host proof stages) or unvalidated=True "
                             "(frame count on assumed boxes, written to plot_parameters_frame_count_only.txt)")
        code_of = frame_det_code
    if path is None:
        path = "plot_parameters_frame_count_only.txt" if unvalidated else "plot_parameters.txt"
    m = len(descs)
    if len(params) != m:
        raise ValueError("SampleParameters: %d descriptors for %d parameter vectors" % (m, len(params)))
    if chunk <= 0:
        raise ValueError("SampleParameters: chunk must be positive")
    jpath = path + ".journal"
    key = _journal_key(descs, m)
    done = _read_journal(jpath, key) if resume else {}
    done = {i: c for i, c in done.items() if 0 <= i < m}
    todo = [i for i in range(m) if i not in done]
    with open(jpath + ".tmp", "w") as jf:              # rewrite what was parsed: a torn last line must not be appended to
        jf.write("# " + key + "\n" + "".join("%d %d\n" % (i, done[i]) for i in sorted(done)))
    os.replace(jpath + ".tmp", jpath)
    with open(jpath, "a") as jf:
        for at in range(0, len(todo), chunk):
            idx = todo[at:at + chunk]
            batch = (abi.FrontDesc * len(idx))()
            for k, i in enumerate(idx):
                C.memmove(C.byref(batch[k]), C.byref(descs[i]), C.sizeof(abi.FrontDesc))
            res = frame_count_batch(batch)
            if len(res) != len(idx):
                raise RuntimeError("SampleParameters: engine returned %d records for %d fronts" % (len(res), len(idx)))
            out = []
            for k, i in enumerate(idx):
                c = code_of(res[k])
                done[i] = c if c >= 0 else -1
                out.append("%d %d\n" % (i, done[i]))
            jf.write("".join(out))
            jf.flush()
            os.fsync(jf.fileno())
    counts = [done[i] for i in range(m)]
    write_plot_parameters(path, params, counts)
    os.remove(jpath)
    return counts
