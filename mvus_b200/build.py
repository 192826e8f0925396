"""Build the CUDA library in-tree: mvus_b200/libmvus_ba.so (sm_100a, -lineinfo).
`python -m mvus_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without a GPU."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'mvus_ba.cu')
OUT = os.path.join(HERE, 'libmvus_ba.so')
DEPS = [os.path.join(HERE, 'csrc', f) for f in os.listdir(os.path.join(HERE, 'csrc'))] + \
       [os.path.join(os.path.dirname(HERE), 'include', 'mvus_ba.h')]


def nvcc_cmd(extra=()):
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    return [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
            '-shared', '-Xcompiler', '-fPIC', '-Xcompiler', '-Wno-unknown-pragmas', '--cudart', 'shared',
            '-o', OUT, SRC, '-ldl', '-lpthread'] + list(extra)


def build(force=False, verbose=False):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in DEPS):
        return OUT
    cmd = nvcc_cmd(['-Xptxas', '-v'] if verbose else [])
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
    if verbose:
        sys.stderr.write(res.stderr)
    return OUT


if __name__ == '__main__':
    print(build(force=True, verbose='-v' in sys.argv))
