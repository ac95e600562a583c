"""tests/emu/build.py -- TEST INFRASTRUCTURE.  Builds the CPU-emulated twin of
libgeepee_b200.so (same kernel + launch source, -DGPB_CPU_EMU, g++ only) into
tests/emu/_build/.  The CPU test-suite uses it to exercise the kernels' logic and the
host code in the GPU-less container; the product never loads it."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..', '..'))
CSRC = os.path.join(ROOT, 'geepee_b200', 'csrc')
OUT = os.path.join(HERE, '_build', 'libgeepee_b200_emu.so')


def build(force=False):
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith('.cu')] + \
        [os.path.join(HERE, 'gpb_emu.cpp')]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.cu'))] + \
        [os.path.join(ROOT, 'include', 'geepee_b200.h')]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    flags = ['-std=c++17', '-O1', '-g', '-fPIC', '-DGPB_CPU_EMU', '-Wall', '-Wno-unused-function',
             '-Wno-unknown-pragmas', '-Wno-unused-variable', '-x', 'c++']
    objs, procs = [], []
    for src in srcs:
        obj = os.path.join(os.path.dirname(OUT), os.path.basename(src) + '.o')
        objs.append(obj)
        procs.append(subprocess.Popen(['g++'] + flags + ['-c', src, '-o', obj]))
    if any(p.wait() != 0 for p in procs):
        raise RuntimeError('emulator build failed')
    subprocess.check_call(['g++', '-shared', '-o', OUT] + objs + ['-lm'])
    return OUT


if __name__ == '__main__':
    print(build(force=True))
