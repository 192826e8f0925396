"""Recipe for oracle/_ref/: a verbatim, UNMODIFIED copy of the reference's pure-Python package
(/root/reference/multiviewunsynch, MPL-2.0) so that the reference itself can be imported on the
GPU box, where /root/reference does not exist.  TEST INFRASTRUCTURE ONLY:

  * oracle/_ref/ is git-ignored (it never enters the history) but NOT gpurun-ignored, so it
    travels with the snapshot exactly like the built .so files;
  * only tests/, __graft_entry__.smoke() and bench.py's CPU arm (`--impl reference`,
    `cpu_baseline`) import it, through oracle/ref_shim.py, as the checker / the CPU baseline;
  * nothing under mvus_b200/ imports it (tests/test_host.py::test_product_never_imports_oracle).

    python oracle/make_ref.py            # no-op when /root/reference is absent
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(os.environ.get('MVUS_REFERENCE', '/root/reference'), 'multiviewunsynch')
DST = os.path.join(HERE, '_ref', 'multiviewunsynch')


def make(verbose=False):
    if not os.path.isdir(os.path.join(SRC, 'reconstruction')):
        if verbose:
            print('reference not present at %s: oracle/_ref left as it is' % SRC)
        return os.path.isdir(os.path.join(DST, 'reconstruction'))
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    if verbose:
        n = sum(len(f) for _, _, f in os.walk(DST))
        print('copied %d files -> %s' % (n, DST))
    return True


if __name__ == '__main__':
    sys.exit(0 if make(verbose=True) else 1)
