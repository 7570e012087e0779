"""TEST INFRASTRUCTURE ONLY -- compile the UNMODIFIED reference CPU operator.

Builds /root/reference/torbi/csrc/{ops.cpp,viterbi.cpp} (read in place, never
copied) into oracle/_ref/torbi_ref_C.so with the flags the reference's own
setup.py uses for its CPU extension (-O3 -fopenmp, setup.py:60-65), plus a tiny
side translation unit oracle/ref_debug_trellis.cpp that only *declares* the
reference's externally-linked torbi::viterbi_make_trellis_cpu
(viterbi.cpp:35) and registers torbi::debug_trellis so tests can see the
reference's fp32 posterior and int32 backpointers.

oracle/_ref/ is git-ignored but NOT gpurun-ignored: the .so built here travels
to the GPU box, where /root/reference does not exist.  `load()` therefore only
needs the prebuilt file.
"""
import os
import pathlib
import sys

HERE = pathlib.Path(__file__).resolve().parent
REF_DIR = HERE / '_ref'
REF_SO = REF_DIR / 'torbi_ref_C.so'
REFERENCE_CSRC = pathlib.Path('/root/reference/torbi/csrc')


def available() -> bool:
    return REF_SO.exists()


def can_build() -> bool:
    return (REFERENCE_CSRC / 'viterbi.cpp').exists()


def build(verbose: bool = False) -> pathlib.Path:
    """Compile the reference sources where they lie; outputs only in oracle/_ref."""
    if not can_build():
        raise FileNotFoundError(f'{REFERENCE_CSRC} not present; cannot build oracle/_ref')
    # the image exports CC/CXX=/opt/gcc/bin/* (a wrapper that cannot find
    # libgomp.spec); the distro compiler on PATH links OpenMP fine
    os.environ['CC'] = '/usr/bin/gcc'
    os.environ['CXX'] = '/usr/bin/g++'
    from torch.utils.cpp_extension import load
    REF_DIR.mkdir(exist_ok=True)
    load(
        name='torbi_ref_C',
        sources=[
            str(REFERENCE_CSRC / 'ops.cpp'),
            str(REFERENCE_CSRC / 'viterbi.cpp'),
            str(HERE / 'ref_debug_trellis.cpp')],
        extra_cflags=['-O3', '-fopenmp'],
        extra_ldflags=['-fopenmp'],
        is_python_module=False,
        build_directory=str(REF_DIR),
        verbose=verbose)
    return REF_SO


REF_CUDA_DIR = REF_DIR / 'cuda'
REF_CUDA_SO = REF_CUDA_DIR / 'torbi_ref_cuda.so'


def cuda_available() -> bool:
    return REF_CUDA_SO.exists()


def build_cuda(verbose: bool = False) -> pathlib.Path:
    """The reference's own CUDA operator (/root/reference/torbi/csrc/cuda/viterbi.cu with ops.cpp),
    UNMODIFIED, cross-compiled here for sm_100a with the reference's flags (-O3, setup.py:70-72) into
    oracle/_ref/cuda/.  Only tools/time_reference_cuda.py loads it (in its own process: it registers
    the same `torbi` operator library as the CPU build) -- it is the kernel-level bar measured
    beside ours on the same B200, never part of the product."""
    if not (REFERENCE_CSRC / 'cuda' / 'viterbi.cu').exists():
        raise FileNotFoundError(f'{REFERENCE_CSRC}/cuda not present; cannot build the reference CUDA op')
    os.environ['CC'] = '/usr/bin/gcc'
    os.environ['CXX'] = '/usr/bin/g++'
    os.environ['TORCH_CUDA_ARCH_LIST'] = '10.0a'
    from torch.utils.cpp_extension import load
    REF_CUDA_DIR.mkdir(parents=True, exist_ok=True)
    load(
        name='torbi_ref_cuda',
        sources=[str(REFERENCE_CSRC / 'ops.cpp'), str(REFERENCE_CSRC / 'cuda' / 'viterbi.cu')],
        extra_cflags=['-O3'],
        extra_cuda_cflags=['-O3'],
        with_cuda=True,
        is_python_module=False,
        build_directory=str(REF_CUDA_DIR),
        verbose=verbose)
    return REF_CUDA_SO


_loaded = False


def load_ref():
    """torch.ops.load_library(oracle/_ref/torbi_ref_C.so); returns torch.ops.torbi."""
    global _loaded
    import torch
    if not _loaded:
        if not available():
            raise FileNotFoundError(
                f'{REF_SO} missing: run `python oracle/build_ref.py` where /root/reference exists')
        torch.ops.load_library(str(REF_SO))
        _loaded = True
    return torch.ops.torbi


if __name__ == '__main__':
    if 'cuda' in sys.argv:
        print(build_cuda(verbose='-v' in sys.argv))
    else:
        print(build(verbose='-v' in sys.argv))
