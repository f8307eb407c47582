"""Build libttg_b200.so in-tree with nvcc for sm_100a (no torch headers, no JIT cache).

    python -m tartangan_b200.build [--force]
"""
import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libttg_b200.so')
# development only: extra -D flags / alternative output (variant builds for A/B kernel timing)
EXTRA_DEFS = os.environ.get('TTG_BUILD_DEFS', '').split()
VARIANT = os.environ.get('TTG_BUILD_VARIANT', '')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isfile(cand) or cand == 'nvcc'):
            return cand
    raise RuntimeError('nvcc not found')


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def is_stale():
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh'))
    return any(os.path.getmtime(p) > t for p in deps)


def _compile_one(src, obj, verbose):
    cmd = [_nvcc()] + NVCC_FLAGS + EXTRA_DEFS + ['-c', src, '-o', obj]
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
    return r.stderr


def build(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a and link the shared library."""
    lib_path = LIB_PATH if not VARIANT else LIB_PATH[:-3] + '_' + VARIANT + '.so'
    if not VARIANT and not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(HERE, 'build' + ('_' + VARIANT if VARIANT else ''))
    os.makedirs(obj_dir, exist_ok=True)
    srcs = sources()
    hdr_t = max([os.path.getmtime(p) for p in glob.glob(os.path.join(CSRC, '*.cuh'))] + [0])
    objs, jobs = [], []
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        for s in srcs:
            o = os.path.join(obj_dir, os.path.basename(s)[:-3] + '.o')
            objs.append(o)
            if force or not os.path.isfile(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_t):
                jobs.append(ex.submit(_compile_one, s, o, verbose))
        for j in jobs:
            log = j.result()
            if verbose and log:
                print(log)
    r = subprocess.run([_nvcc(), '-shared', '-o', lib_path] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    return lib_path


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
