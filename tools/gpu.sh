#!/bin/bash
# build in-tree, then run a command on the GPU box (only if the build succeeded)
cd /root/repo || exit 1
python -c "import __graft_entry__ as g; g.build()" 2>&1 | grep -i "error\|fail" && exit 1
timeout 3000 /usr/local/graft/bin/gpurun --timeout ${GPU_TIMEOUT:-900} -- "$1" 2>&1 | tail -${TAIL:-12}
