"""Spatial partition of ONE GPU between the actor replicas and the learner (CUDA green contexts, driver API >= 12.4).

In the reference's a0-l0 topology the actor threads and the learner share GPU 0 and run concurrently
(cleanba/cleanba_ppo.py:669-686: rollout threads are started next to the learner loop).  On a B200 the learner's
persistent tcgen05 kernels occupy every SM, so an actor kernel launched beside them waits for a whole learner kernel
(~0.1-0.3 ms) at every one of its ~22 launches.  A green context gives the actor streams a small private set of SMs
(multiples of 8 on sm_100) and the learner the rest; kernels of the two sides then never queue behind each other, and the
launch-latency-bound actor step (n = 60) runs in the shadow of the HBM-bound learner step.

Only kernel launches are confined to a partition; copies use the copy engines as usual.  Contexts that launch into a partition
must size their persistent grids for it: `Context.set_sm_budget(partition.actor_sms | learner_sms)`.
"""
import torch

from .lib import CleanbaError


def _chk(ret):
    from cuda.bindings import driver as drv
    err = ret[0]
    if err != drv.CUresult.CUDA_SUCCESS:
        raise CleanbaError(f"CUDA driver error {err!r} while partitioning the GPU")
    return ret[1:] if len(ret) > 2 else ret[1]


class SmPartition:
    def __init__(self, device, actor_sms: int = 16):
        try:
            from cuda.bindings import driver as drv
        except Exception as e:  # pragma: no cover
            raise CleanbaError(f"cuda-python (cuda.bindings.driver) is needed for SM partitions: {e}")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise CleanbaError("SM partitions live on CUDA devices only")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        with torch.cuda.device(self.device):
            torch.zeros(1, device=self.device)            # primary context
            dev = _chk(drv.cuDeviceGet(idx))
            sm = _chk(drv.cuDeviceGetDevResource(dev, drv.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
            groups, nb, remaining = _chk(drv.cuDevSmResourceSplitByCount(1, sm, 0, actor_sms))
            if nb < 1:
                raise CleanbaError(f"cannot split off {actor_sms} SMs")
            self.actor_sms = int(groups[0].sm.smCount)
            self.learner_sms = int(remaining.sm.smCount)
            self._desc_a = _chk(drv.cuDevResourceGenerateDesc([groups[0]], 1))
            self._desc_l = _chk(drv.cuDevResourceGenerateDesc([remaining], 1))
            flag = drv.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM
            self._gctx_a = _chk(drv.cuGreenCtxCreate(self._desc_a, dev, flag))
            self._gctx_l = _chk(drv.cuGreenCtxCreate(self._desc_l, dev, flag))
        self._drv = drv
        self._streams = []

    def _stream(self, gctx, priority):
        drv = self._drv
        with torch.cuda.device(self.device):
            s = _chk(drv.cuGreenCtxStreamCreate(gctx, drv.CUstream_flags.CU_STREAM_NON_BLOCKING, priority))
        self._streams.append(s)
        return torch.cuda.ExternalStream(int(s), device=self.device)

    def actor_stream(self) -> "torch.cuda.ExternalStream":
        """A new stream whose kernels run on the actor partition."""
        return self._stream(self._gctx_a, -1)

    def learner_stream(self) -> "torch.cuda.ExternalStream":
        """A new stream whose kernels run on the learner partition."""
        return self._stream(self._gctx_l, 0)
