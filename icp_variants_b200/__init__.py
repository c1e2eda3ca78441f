"""icp_variants_b200 -- the host-side mirror of ICP-Variants' registration path over the icp_gpu_* C ABI (include/icp_gpu.h).

    capi        ctypes binding of lib/libicp_gpu.so (the sm_100a kernels); raises when the library is missing -- there is no CPU path
    optimizer   LinearICPOptimizer / CeresICPOptimizer with the reference's setters (ICPOptimizer.h:41-95), ConvergenceMeasure, TimeMeasure
    sequence    reconstructRoom (main.cpp:183-341) and alignPairs (the ETH pair loop, main.cpp:411-498) on resident contexts
    experiment  the CSV experiment runner (experiment.cpp)
    parallel    pair queues per rank, point shards and the peer-mailbox attachment for torch.distributed launches
    io          OFF / PCD / PLY / point-cloud dump readers and writers, TUM lists, ETH CSV
    synth       synthetic clouds in the shapes of the reference's data sets (bunny, TUM frames, ETH Apartment scans)

Nothing here imports oracle/ (test infrastructure)."""
__version__ = "0.2.0"
__all__ = ["capi", "optimizer", "sequence", "experiment", "parallel", "io", "synth"]


def __getattr__(name):
    # submodules on first use: importing the package must not pull numpy-heavy modules or open the CUDA library
    if name in __all__:
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
