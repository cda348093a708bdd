"""Host BLAS thread control for the small factorisations on the host side of the path (L-BFGS steps, the M x M solves of
Thompson sampling): on a many-core host a multi-threaded BLAS spends ~80 ms per call spinning on 100 x 100 matrices, one
thread 0.3 ms.  The controller is built once: constructing it scans every loaded shared library (~1 ms)."""
import contextlib

_controller = None


def single_threaded_blas():
    """Context manager: BLAS / OpenMP pools limited to one thread (no-op without threadpoolctl)."""
    global _controller
    try:
        from threadpoolctl import ThreadpoolController
    except ImportError:  # pragma: no cover
        return contextlib.nullcontext()
    if _controller is None:
        _controller = ThreadpoolController()
    return _controller.limit(limits=1)
