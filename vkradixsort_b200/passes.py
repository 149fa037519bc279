"""Python mirror of the reference's pass objects, driven through the C-ABI.

This is a *test / bench driver*: the native drop-in is the C++ facade in
include/vkradixsort_b200.hpp, which carries the same names.  Names, argument meaning and error
behaviour follow the reference so the parity tests read like the reference's own examples:

  MultiRadixSortPass   multiradixsort/include/MultiRadixSortPass.h:7-40, src/MultiRadixSortPass.cpp:10-20
  SingleRadixSortPass  singleradixsort/include/SingleRadixSortPass.h:7-28
  ComputePass methods  engine/include/engine/passes/ComputePass.h:16-60
  setStorageBuffer     engine/include/engine/passes/Pass.h:54-104
  MultiRadixSort / SingleRadixSort (the 4-iteration loop and the ping-pong bindings)
                       multiradixsort/src/MultiRadixSort.cpp:12-62, singleradixsort/src/SingleRadixSort.cpp:12-28

Buffers are torch CUDA tensors (device memory + streams are torch's job here, nothing else).
"""
from __future__ import annotations

from . import capi

MAX_FRAMES_IN_FLIGHT = 2  # engine/include/engine/core/GPUContext.h:110


class GPUContext:
    """The two things the sort needs from engine::GPUContext: a device and the ping-pong frame
    counter (getActiveIndex / incrementActiveIndex, GPUContext.h:71-73)."""

    def __init__(self, device: int = 0, stream=None):
        self.device = device
        self.stream = stream
        self._active_index = 0

    def getActiveIndex(self) -> int:
        return self._active_index

    def incrementActiveIndex(self) -> None:
        self._active_index = (self._active_index + 1) % MAX_FRAMES_IN_FLIGHT

    def getMultiBufferedCount(self) -> int:
        return MAX_FRAMES_IN_FLIGHT


class _ComputePass:
    """ComputePass + the parts of Pass a caller touches."""

    _local_size = capi.WORKGROUP_SIZE  # every shader: layout(local_size_x = 256)
    _num_stages = 1

    def __init__(self, gpuContext: GPUContext):
        self.m_gpuContext = gpuContext
        self._handle: capi.Handle | None = None
        self._work_group_counts = [(0, 0, 0)] * self._num_stages
        # bindings[frame][(set, binding)] = buffer
        self._bindings = [dict() for _ in range(MAX_FRAMES_IN_FLIGHT)]

    def create(self, max_num_elements_hint: int = 0) -> None:
        self._handle = capi.Handle(self.m_gpuContext.device, max_num_elements_hint)

    def release(self) -> None:
        if self._handle is not None:
            self._handle.close()
            self._handle = None

    def setGlobalInvocationSize(self, stageIndex: int, width: int, height: int, depth: int) -> None:
        # ComputePass::getDispatchSize, ComputePass.h:24-29 (local size is (256,1,1))
        x = capi.workgroup_count(width)
        self._work_group_counts[stageIndex] = (x, height, depth)

    def getWorkGroupCount(self, stageIndex: int):
        class _Extent:
            def __init__(self, w, h, d):
                self.width, self.height, self.depth = w, h, d

        return _Extent(*self._work_group_counts[stageIndex])

    def setStorageBuffer(self, *args) -> None:
        """setStorageBuffer(set, binding, buffer) binds for all frames;
        setStorageBuffer(frame, set, binding, buffer) for one (Pass.h:54-104)."""
        if len(args) == 3:
            set_, binding, buf = args
            for f in range(MAX_FRAMES_IN_FLIGHT):
                self._bindings[f][(set_, binding)] = buf
        elif len(args) == 4:
            frame, set_, binding, buf = args
            self._bindings[frame][(set_, binding)] = buf
        else:
            raise TypeError("setStorageBuffer(set, binding, buffer) or (frame, set, binding, buffer)")

    def _bound(self, set_: int, binding: int):
        frame = self.m_gpuContext.getActiveIndex()
        try:
            return self._bindings[frame][(set_, binding)]
        except KeyError:
            # the reference asserts on a missing binding (Pass.h:59-60)
            raise AssertionError(f"no storage buffer bound at frame {frame}, set {set_}, binding {binding}")

    def _require(self) -> capi.Handle:
        if self._handle is None:
            raise RuntimeError("pass not created")  # reference: using a pass before create() is UB
        return self._handle


class MultiRadixSortPass(_ComputePass):
    RADIX_SORT_HISTOGRAMS = 0  # ComputeStage, MultiRadixSortPass.h:12-15 (also the descriptor set number)
    RADIX_SORT = 1
    _num_stages = 2

    PushConstantsHistograms = capi.MultiPushConstants
    PushConstants = capi.MultiPushConstants

    def __init__(self, gpuContext: GPUContext):
        super().__init__(gpuContext)
        self.m_pushConstantsHistogram = capi.MultiPushConstants()
        self.m_pushConstants = capi.MultiPushConstants()

    def execute(self, awaitBeforeExecution=None):
        """One pass for the current g_shift = recordCommands (MultiRadixSortPass.cpp:10-20):
        stage 0 on (0,0)->(0,1), barrier, stage 1 on (1,0),(1,2)->(1,1), barrier.  Stream order
        replaces the semaphores; the return value only keeps the chaining idiom alive."""
        h = self._require()
        stream = self.m_gpuContext.stream
        h.multi_histograms(self._bound(0, 0), self._bound(0, 1), self.m_pushConstantsHistogram, stream=stream)
        h.multi_scatter(self._bound(1, 0), self._bound(1, 1), self._bound(1, 2), self.m_pushConstants, stream=stream)
        return object()


class SingleRadixSortPass(_ComputePass):
    RADIX_SORT = 0  # SingleRadixSortPass.h:12-14
    _num_stages = 1
    PushConstants = capi.SinglePushConstants

    def __init__(self, gpuContext: GPUContext):
        super().__init__(gpuContext)
        self.m_pushConstants = capi.SinglePushConstants()

    def execute(self, awaitBeforeExecution=None):
        h = self._require()
        h.single_sort(self._bound(0, 0), self._bound(0, 1), self.m_pushConstants, stream=self.m_gpuContext.stream)
        return object()


class MultiRadixSort:
    """Program logic of MultiRadixSort::execute (MultiRadixSort.cpp:5-81) without the RNG /
    printing: sizing, push constants, ping-pong bindings, the 4-iteration loop."""

    NUM_BLOCKS_PER_WORKGROUP = 32  # MultiRadixSort.cpp:12
    NUM_ITERATIONS = 4             # MultiRadixSort.cpp:51-55 (SORT_32BIT)

    def __init__(self, nb: int | None = None):
        if nb is not None:
            self.NUM_BLOCKS_PER_WORKGROUP = nb
        self.m_pass: MultiRadixSortPass | None = None

    def histogram_buffer_elements(self, num_elements: int) -> int:
        gis = capi.global_invocation_size(num_elements, self.NUM_BLOCKS_PER_WORKGROUP)
        return max(1, capi.workgroup_count(gis)) * capi.RADIX_SORT_BINS  # MultiRadixSort.cpp:93

    def execute(self, gpuContext: GPUContext, buffers, num_elements: int) -> None:
        """buffers = [keys in/out, scratch, histograms] (m_buffers, MultiRadixSort.h:33).
        Enqueues the whole staged sort; the caller synchronises (vkQueueWaitIdle, :62)."""
        P = MultiRadixSortPass
        self.m_pass = p = MultiRadixSortPass(gpuContext)
        p.create(num_elements)
        try:
            nb = self.NUM_BLOCKS_PER_WORKGROUP
            gis = capi.global_invocation_size(num_elements, nb)  # :13-15
            p.setGlobalInvocationSize(P.RADIX_SORT_HISTOGRAMS, gis, 1, 1)
            p.setGlobalInvocationSize(P.RADIX_SORT, gis, 1, 1)
            W = p.getWorkGroupCount(P.RADIX_SORT_HISTOGRAMS).width
            assert W == p.getWorkGroupCount(P.RADIX_SORT).width  # :21
            for pc in (p.m_pushConstantsHistogram, p.m_pushConstants):  # :22-27
                pc.g_num_elements = num_elements
                pc.g_num_workgroups = W
                pc.g_num_blocks_per_workgroup = nb
            a = gpuContext.getActiveIndex()  # :34-46
            b = (a + 1) % 2
            p.setStorageBuffer(a, P.RADIX_SORT_HISTOGRAMS, 0, buffers[0])
            p.setStorageBuffer(a, P.RADIX_SORT, 0, buffers[0])
            p.setStorageBuffer(b, P.RADIX_SORT, 1, buffers[0])
            p.setStorageBuffer(b, P.RADIX_SORT_HISTOGRAMS, 0, buffers[1])
            p.setStorageBuffer(a, P.RADIX_SORT, 1, buffers[1])
            p.setStorageBuffer(b, P.RADIX_SORT, 0, buffers[1])
            p.setStorageBuffer(P.RADIX_SORT_HISTOGRAMS, 1, buffers[2])
            p.setStorageBuffer(P.RADIX_SORT, 2, buffers[2])
            await_ = None
            for i in range(self.NUM_ITERATIONS):  # :56-61
                p.m_pushConstantsHistogram.g_shift = 8 * i
                p.m_pushConstants.g_shift = 8 * i
                await_ = p.execute(await_)
                gpuContext.incrementActiveIndex()
        finally:
            # release() synchronises the device, so the enqueued passes are complete on return
            p.release()


class SingleRadixSort:
    """SingleRadixSort::execute (SingleRadixSort.cpp:5-47): one work group, one dispatch."""

    def execute(self, gpuContext: GPUContext, buffers, num_elements: int) -> None:
        P = SingleRadixSortPass
        p = SingleRadixSortPass(gpuContext)
        p.create(num_elements)
        try:
            p.setGlobalInvocationSize(P.RADIX_SORT, 256, 1, 1)  # :12
            p.m_pushConstants.g_num_elements = num_elements     # :15
            p.setStorageBuffer(P.RADIX_SORT, 0, buffers[0])      # :22-23
            p.setStorageBuffer(P.RADIX_SORT, 1, buffers[1])
            p.execute(None)
        finally:
            p.release()
