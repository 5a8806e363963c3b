"""In-tree build of libsedb200.so (nvcc, sm_100a only).

    python -m sound_event_detection_dcase2017_task4_b200.build [--force] [--verbose]

The shared object lands next to this file so that it travels with the repo snapshot to the
GPU box (it is git-ignored, not gpurun-ignored).  nvcc cross-compiles without a GPU.
"""
import concurrent.futures
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, 'csrc')
OBJ_DIR = os.path.join(PKG_DIR, 'csrc', 'build')
LIB_PATH = os.path.join(PKG_DIR, 'libsedb200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a']
NVCC_FLAGS = ['-O3', '-std=c++17', '-lineinfo', '--use_fast_math=false', '-Xcompiler', '-fPIC',
              '-Xcompiler', '-fvisibility=hidden', '-Xptxas', '-v', '--expt-relaxed-constexpr']
NVCC_FLAGS = [f for f in NVCC_FLAGS if f != '--use_fast_math=false']


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    hs.append(os.path.join(os.path.dirname(PKG_DIR), 'include', 'sed_b200.h'))
    return hs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, force, verbose):
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + '.o')
    if not force and not _stale(obj, [src] + headers()):
        return obj, ''
    cmd = [NVCC] + ARCH_FLAGS + NVCC_FLAGS + ['-c', src, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
    return obj, r.stderr if verbose else ''


def build_variant(tag, defines):
    """Development: a second library libsedb200_<tag>.so compiled with extra -D switches (A/B runs under gpurun select
    it with SED_B200_LIB=<path>, see _lib.py).  Not used by the product."""
    global OBJ_DIR, LIB_PATH, NVCC_FLAGS
    saved = (OBJ_DIR, LIB_PATH, NVCC_FLAGS)
    try:
        OBJ_DIR = os.path.join(PKG_DIR, 'csrc', 'build_' + tag)
        LIB_PATH = os.path.join(PKG_DIR, 'libsedb200_%s.so' % tag)
        NVCC_FLAGS = NVCC_FLAGS + ['-D' + d for d in defines]
        return build(force=False)
    finally:
        OBJ_DIR, LIB_PATH, NVCC_FLAGS = saved


def build(force=False, verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                sys.stderr.write(log)
    if force or _stale(LIB_PATH, objs):
        cmd = [NVCC] + ARCH_FLAGS + ['-shared', '-o', LIB_PATH] + objs + ['-cudart', 'static']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    return LIB_PATH


if __name__ == '__main__':
    if '--variant' in sys.argv:
        i = sys.argv.index('--variant')
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
