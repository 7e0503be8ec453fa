"""One process per GPU: wiring the engines' per-round exchange over torch.distributed.

PyTorch is plumbing here (process group, symmetric memory allocation, handle exchange); the exchange
itself runs inside the persistent kernel (csrc/ts_persist.cuh) as NVLink stores or NVLS multimem
operations.  Reference counterpart: the main thread adding the workers' lambdat partial sums
(src/snpsamplinge.cc:337-352)."""
import os

from . import capi

_keepalive = []  # symmetric buffers must outlive the engines that use them

MODES = {3: "the CTA whose arrival completes a word (atomic with return value) adds the GPU's total into an accumulator on "
            "every rank (one NVLink red.add.u64 per peer), every CTA polls one local word pair",
         4: "the CTA whose arrival completes a word (atomic with return value) adds the GPU's total into every rank's "
            "accumulator with one multimem.red.add.u64 (NVLS), every CTA polls one local word pair"}


def connect(engine, group=None):
    """Connect `engine` (rank, nranks as created) with its peers in `group` (default: WORLD).

    Default: a torch symmetric-memory buffer with its NVLS multicast alias (ts_comm_attach_symmetric): on every
    GPU the CTA whose arrival completes a word adds the GPU's total into an accumulator on every rank with ONE
    multimem.red and every CTA polls one local word pair (XMODE_MCACC, the fastest scheme measured:
    profiles/r2_summary.md).  Without symmetric memory / NVLS, or with TSGPU_XCHG=gacc: CUDA-IPC handles of the
    engines' own buffers (ts_comm_export / ts_comm_connect) and one NVLink red.add per peer (XMODE_GACC).
    Returns a dict describing what was set up; ends with a barrier, so ts_steps may follow immediately."""
    import torch
    import torch.distributed as dist
    group = group or dist.group.WORLD
    world = dist.get_world_size(group)
    info = {"exchange": "none (one rank)", "detail": ""}
    if world == 1:
        return info
    want = os.environ.get("TSGPU_XCHG", "")
    if want not in ("gacc", "slots", "ipc"):  # "slots" / "ipc": names of earlier builds, now the same as "gacc"
        import torch.distributed._symmetric_memory as symm_mem
        dev = torch.device("cuda", int(engine.cfg.device))
        nbytes = int(capi.lib().ts_comm_state_bytes())
        got, why = None, ""
        try:
            buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=dev)
            hdl = symm_mem.rendezvous(buf, group)
            off = int(getattr(hdl, "offset", 0) or 0)
            ptrs = [int(p) + off for p in hdl.buffer_ptrs]
            if ptrs[dist.get_rank(group)] != buf.data_ptr():
                raise RuntimeError("unexpected symmetric-memory layout")
            mc = int(hdl.multicast_ptr or 0)
            got = (buf, hdl, ptrs, mc + off if mc else 0)
        except Exception as ex:  # no symmetric memory on this fabric / build
            why = "%s: %s" % (type(ex).__name__, str(ex)[:200])
        # every rank must take the same branch
        ok = torch.tensor([1 if got else 0, 1 if (got and got[3]) else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok[0]):
            buf, hdl, ptrs, mc = got
            if not int(ok[1]):
                mc = 0
            engine.attach_symmetric(ptrs, mc, nbytes)
            _keepalive.append((buf, hdl))
            torch.cuda.synchronize(dev)
            dist.barrier(group)
            return {"exchange": MODES[engine.comm_mode], "mode": engine.comm_mode, "multicast": bool(mc),
                    "detail": "torch symmetric memory"}
        info["detail"] = "symmetric memory unavailable (%s); " % why
    handles = [None] * world
    dist.all_gather_object(handles, engine.comm_export(), group=group)
    engine.comm_connect(handles)
    dist.barrier(group)
    info["exchange"] = MODES[engine.comm_mode]
    info["mode"] = engine.comm_mode
    info["multicast"] = False
    info["detail"] += "CUDA-IPC peer mappings"
    return info
