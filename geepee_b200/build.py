"""Build libgeepee_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(CSRC, 'libgeepee_b200.so')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared']


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith('.cu')]


def needs_build():
    if not os.path.exists(OUT):
        return True
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh'))]
    deps.append(os.path.join(HERE, '..', 'include', 'geepee_b200.h'))
    return any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps)


def build(force=False, verbose=False, out=None, extra_flags=()):
    """nvcc -c every .cu in parallel (one process per translation unit), then link.
    `out` / `extra_flags`: development variants of the library (A/B builds for tools/kbench.py,
    selected at run time with GPB_LIB_PATH); the product is the default build."""
    variant = out is not None
    if not variant and not force and not needs_build():
        return OUT
    out = out or OUT
    nvcc = os.environ.get('NVCC', 'nvcc')
    objdir = os.path.join(CSRC, 'build' if not variant else 'build_' + os.path.basename(out))
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != '-shared'] + (['-Xptxas', '-v'] if verbose else []) + \
        os.environ.get('GPB_EXTRA_NVCC_FLAGS', '').split() + list(extra_flags)
    procs = []
    objs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        log = open(obj + '.log', 'w')
        procs.append((src, log, subprocess.Popen([nvcc] + flags + ['-c', src, '-o', obj],
                                                 stdout=log, stderr=subprocess.STDOUT)))
    for src, log, p in procs:
        rc = p.wait()
        log.close()
        if rc != 0:
            raise RuntimeError('nvcc failed on %s:\n%s' % (src, open(log.name).read()[-4000:]))
    subprocess.check_call([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', out] + objs)
    return out


if __name__ == '__main__':
    import sys
    if len(sys.argv) > 2:       # python -m geepee_b200.build <out.so> <nvcc flags...>
        print(build(force=True, out=os.path.abspath(sys.argv[1]), extra_flags=sys.argv[2:]))
    else:
        print(build(force=True))
