#!/bin/bash
# Puts the UNMODIFIED reference (shinmorino/qgate 0.2.2) under baseline/_ref/ — git-ignored, but it
# travels to the GPU box with the gpurun snapshot — so that the reference's own Python front end and
# its own unittest suite can run there on top of libqgate_b200.so (tests/test_reference_frontend.py,
# bench.py's e2e leg).  Run in the build container, where /root/reference exists.
#
# `pip install --target baseline/_ref /root/reference` is not usable: the reference's setup.py runs
# `make cuda_obj` for sm_35 ... sm_70, which nvcc 12.9 rejects.  Instead: scratch copy, its own
# Makefile targets for glue.so + cpuext.so (the recipe of SURVEY.md section 8c), then the package,
# its tests and its examples are copied as they are.  Nothing lands in tracked files.
set -e
REF=${REF:-/root/reference}
SCRATCH=${SCRATCH:-/tmp/qgate_ref}
HERE=$(cd "$(dirname "$0")/.." && pwd)
if [ ! -d "$REF/qgate" ]; then echo "no reference at $REF"; exit 0; fi
if [ ! -f "$SCRATCH/qgate/simulator/cpuext.so" ]; then
  rm -rf "$SCRATCH" && cp -r "$REF" "$SCRATCH" && chmod -R u+w "$SCRATCH"
  (cd "$SCRATCH/qgate/simulator/src" && python3 incpathgen.py > incpath && make ../glue.so ../cpuext.so -j8 > /dev/null 2>&1)
fi
rm -rf "$HERE/baseline/_ref" && mkdir -p "$HERE/baseline/_ref"
cp -r "$SCRATCH/qgate" "$SCRATCH/tests" "$SCRATCH/examples" "$HERE/baseline/_ref/"
rm -rf "$HERE/baseline/_ref/qgate/simulator/src"           # sources stay where they are; only what runs is needed
find "$HERE/baseline/_ref" -name __pycache__ -prune -exec rm -rf {} +
echo "installed: $(du -sh "$HERE/baseline/_ref" | cut -f1) in baseline/_ref"
