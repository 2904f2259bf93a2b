#!/bin/bash
# gpurun for perf probes that do not need the exported reference weights (43 MB under weights_ref):
# the snapshot is pushed without them (seeded random weights are used instead), then .gpurunignore is restored.
# usage: tools/gpurun_lite.sh [--timeout S] -- '<command>'
cd "$(dirname "$0")/.." || exit 1
cp -f .gpurunignore .gpurunignore.bak 2>/dev/null
echo "weights_ref" >> .gpurunignore
/usr/local/graft/bin/gpurun "$@"
rc=$?
if [ -f .gpurunignore.bak ]; then mv -f .gpurunignore.bak .gpurunignore; else rm -f .gpurunignore; fi
exit $rc
