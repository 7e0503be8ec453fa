"""One process per GPU: wiring the engines' per-round exchange over torch.distributed.

PyTorch is plumbing here (process group, symmetric memory allocation, handle exchange); the exchange
itself runs inside the persistent kernel (csrc/ts_persist.cuh) as NVLink stores or NVLS multimem
operations.  Reference counterpart: the main thread adding the workers' lambdat partial sums
(src/snpsamplinge.cc:337-352)."""
import os

from . import capi

_keepalive = []  # symmetric buffers must outlive the engines that use them


def connect(engine, group=None):
    """Connect `engine` (rank, nranks as created) with its peers in `group` (default: WORLD).

    Default: CUDA-IPC handles of the engines' own exchange buffers and peer stores over NVLink
    (ts_comm_export / ts_comm_connect) -- the fastest scheme measured (profiles/r2_summary.md).
    TSGPU_XCHG=slots|mcslot|mcred: a torch symmetric-memory buffer, optionally with its NVLS multicast
    alias (ts_comm_attach_symmetric): one multicast store of the GPU totals, or the in-switch reduction
    with multimem.red from every CTA.  Returns a dict describing what was set up; ends with a barrier,
    so ts_steps may follow immediately."""
    import torch
    import torch.distributed as dist
    group = group or dist.group.WORLD
    world = dist.get_world_size(group)
    info = {"exchange": "none (one rank)", "detail": ""}
    if world == 1:
        return info
    want = os.environ.get("TSGPU_XCHG", "")
    if want in ("slots", "mcslot", "mcred"):
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
            mode = engine.comm_mode
            return {"exchange": {0: "peer stores into symmetric-memory slots (CTA 0 forwards the GPU totals)",
                                 1: "NVLS multicast store of the GPU totals (multimem.st), slots polled locally",
                                 2: "NVLS in-switch reduction: multimem.red.add.u64 from every CTA of every GPU"}[mode],
                    "mode": mode, "multicast": bool(mc), "detail": "torch symmetric memory"}
        info["detail"] = "symmetric memory unavailable (%s); " % why
    handles = [None] * world
    dist.all_gather_object(handles, engine.comm_export(), group=group)
    engine.comm_connect(handles)
    dist.barrier(group)
    info["exchange"] = "peer stores into CUDA-IPC slots (CTA 0 forwards the GPU totals)"
    info["mode"] = 0
    info["multicast"] = False
    return info
