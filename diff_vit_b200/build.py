"""Builds libp2vit_b200.so (the C-ABI library with every sm_100a kernel) in-tree with nvcc.

    python -m diff_vit_b200.build [--force]

The library has no PyTorch dependency: plain `extern "C"` entry points declared in include/p2v.h.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libp2vit_b200.so')
SOURCES = ['p2v_engine.cu', 'p2v_gemm.cu', 'p2v_rowops.cu', 'p2v_attention.cu', 'p2v_attention_tc.cu', 'p2v_observe.cu',
           'p2v_modules.cu', 'p2v_swin.cu']
# -fmad=false: every fused multiply-add of these kernels is written explicitly (__fmaf_rn / __ffma2_rn); the compiler
# must not contract the separately rounded mul + add pairs of the reference's op order (it does contract the packed
# __fmul2_rn + __fadd2_rn intrinsics otherwise)
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-fmad=false',
              '--expt-relaxed-constexpr', '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, '..', 'include', 'p2v.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, 'build', src.replace('.cu', '.o'))
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append('== %s ==\n%s' % (src, out))
        if p.returncode != 0:
            raise RuntimeError('nvcc failed on %s:\n%s' % (src, out))
    subprocess.check_call([nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'])
    with open(os.path.join(HERE, 'build', 'ptxas.log'), 'w') as f:
        f.write('\n'.join(log))
    if verbose:
        print('\n'.join(log))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
