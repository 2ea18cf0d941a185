"""Host mirror of `LoadingManager` (/root/reference/src/app/scene/sdf/loading.rs): the
coarse-to-fine interlaced visit order of the grid fill.  On the GPU a whole pass is one kernel
launch, so this iterator is used for progress bookkeeping, for mapping a pass to its lattice
(`pass_lattice`) and by the tests that port the reference's own unit tests (loading.rs:117-171)."""


def prev_power_of_2(x):  # loading.rs:108-115 (u32)
    x &= 0xFFFFFFFF
    x |= x >> 1
    x |= x >> 2
    x |= x >> 4
    x |= x >> 8
    x |= x >> 16
    return x - (x >> 1)


class LoadingManager:
    def __init__(self, limits, passes):  # loading.rs:23-34
        self.limits = tuple(int(v) for v in limits)
        self.reset(passes)

    def reset(self, passes):  # :37-43
        self.passes = int(passes)
        self.step_size = 2 ** (max(self.passes, 1) - 1)
        self.next_index = [0, 0, 0]
        self.iterations = 0
        self._total_iterations = 0

    def __iter__(self):
        return self

    def __next__(self):  # :50-76
        if self.step_size == 0:
            raise StopIteration
        self.iterations += 1
        self._total_iterations += 1
        res = tuple(self.next_index)
        self.next_index[0] += self.step_size
        if self.next_index[0] >= self.limits[0]:
            self.next_index[0] = 0
            self.next_index[1] += self.step_size
            if self.next_index[1] >= self.limits[1]:
                self.next_index[1] = 0
                self.next_index[2] += self.step_size
                if self.next_index[2] >= self.limits[2]:
                    self.step_size = prev_power_of_2(self.step_size - 1)
                    self.next_index = [0, 0, 0]
                    self.iterations = 0
        return res

    def __len__(self):  # :80-89
        step, it = self.step_size, 0
        while step > 0:
            it += pass_items(self.limits, step)
            step = prev_power_of_2(step - 1)
        return it - self.iterations

    def total_iterations(self):  # :94-96
        return self._total_iterations

    def passes_left(self):  # :99-105
        if self.step_size == 0:
            return 0
        return int(self.step_size).bit_length()  # log2(step) + 1 for a power of two


def pass_items(limits, step):
    """Iterations of one pass: prod ceil(limit / step) (loading.rs:84-85)."""
    n = 1
    for lim in limits:
        n *= (lim + step - 1) // step
    return n


def pass_steps(passes):
    """Step sizes of the passes a fresh manager runs, coarse to fine."""
    step, out = 2 ** (max(int(passes), 1) - 1), []
    while step > 0:
        out.append(step)
        step = prev_power_of_2(step - 1)
    return out


class NativeLoadingManager:
    """The library's own LoadingManager object (`sdfgpu_loading_*`, include/sdfgpu.h): the counters and
    cursor every handle keeps, device-free.  Same interface as `LoadingManager` plus `next_run`."""

    def __init__(self, limits, passes):
        import ctypes as C
        from . import _lib
        self._C, self._lib = C, _lib.load()
        self.limits = tuple(int(v) for v in limits)
        self.passes = int(passes)
        h = C.c_void_p()
        _lib.check(self._lib.sdfgpu_loading_create((C.c_uint32 * 3)(*self.limits), self.passes, C.byref(h)))
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.sdfgpu_loading_destroy(self._h)
            self._h = None

    def reset(self, passes):
        self.passes = int(passes)
        self._lib.sdfgpu_loading_reset(self._h, self.passes)

    def __iter__(self):
        return self

    def __next__(self):
        out = (self._C.c_uint32 * 3)()
        if not self._lib.sdfgpu_loading_next(self._h, out):
            raise StopIteration
        return tuple(out)

    def next_run(self, max_iters):
        """(first_index, step, count): up to max_iters consecutive iterations inside one x row."""
        first, step = (self._C.c_uint32 * 3)(), self._C.c_uint32()
        n = self._lib.sdfgpu_loading_next_run(self._h, int(max_iters), first, self._C.byref(step))
        return (tuple(first), step.value, n) if n else None

    def __len__(self):
        return self._lib.sdfgpu_loading_len(self._h)

    def total_iterations(self):
        return self._lib.sdfgpu_loading_total_iterations(self._h)

    def passes_left(self):
        return self._lib.sdfgpu_loading_passes_left(self._h)
