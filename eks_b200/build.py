"""Build libeks_b200.so (sm_100a) in-tree with nvcc.  Usage: python -m eks_b200.build [--force]"""

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libeks_b200.so')
SOURCES = ["generic.cu", "generic_runs.cu", "lin_lag.cu", "ensemble.cu", "prestage.cu", "diag.cu", "diag_lag.cu", "diag_smooth.cu", "epilogue.cu", "triangulate.cu"]
NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
    '--expt-relaxed-constexpr', '-Xcompiler', '-fPIC', '-Xcompiler', '-O2',
    # the image exports CC/CXX=/opt/gcc/bin/*; nvcc needs the distro host compiler
    '-ccbin', '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++',
]


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError('nvcc not found')


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, '..', 'include', 'eks_b200.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def _obj_stale(src_path: str, obj: str) -> bool:
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    hdrs.append(os.path.join(HERE, '..', 'include', 'eks_b200.h'))
    return any(os.path.getmtime(d) > t for d in [src_path, *hdrs])


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile stale translation units (in parallel) and relink libeks_b200.so."""
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        src_path = os.path.join(CSRC, src)
        obj = os.path.join(LIB_DIR, src.replace('.cu', '.o'))
        objs.append(obj)
        if not force and not _obj_stale(src_path, obj):
            continue
        cmd = [_nvcc(), *NVCC_FLAGS, *os.environ.get('EKS_NVCC_DEFS', '').split(), '-c', src_path, '-o', obj]
        if verbose:
            cmd.insert(1, '-Xptxas=-v')
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}')
    link = [_nvcc(), '-shared', '-o', LIB_PATH, *objs, '-ccbin',
            '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++']
    subprocess.run(link, check=True)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
