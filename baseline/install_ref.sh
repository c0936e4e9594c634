#!/bin/sh
# Installs the UNMODIFIED reference (LienM/recpack at /root/reference) into baseline/_ref so that it travels to
# the GPU box with the repo snapshot (baseline/_ref is git-ignored, not gpurun-ignored).  It is pure Python.
#   - /root/reference is read-only and the build writes into the source tree -> install from a copy in /tmp;
#   - the reference pins numpy/scipy/scikit-learn "==1.*" (setup.py:15-17), this image has numpy 2.3 -> --no-deps
#     (the hot path runs unchanged on the installed versions).
# `hyperopt` (imported by recpack/pipelines/pipeline.py:13, absent here) is provided by baseline/stubs/.
set -e
cd "$(dirname "$0")/.."
rm -rf /tmp/recpack_refsrc baseline/_ref
cp -r /root/reference /tmp/recpack_refsrc
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target baseline/_ref /tmp/recpack_refsrc
rm -rf /tmp/recpack_refsrc
