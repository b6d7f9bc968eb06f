#!/usr/bin/env bash
# gpurun with retries while the pod has no free GPU slot (status=transient, nothing charged).
#   tools/gpurun_retry.sh <logfile> <gpurun args...>
log="$1"; shift
for i in $(seq 1 40); do
    /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
    rc=$?
    if grep -q "status=transient" "$log"; then sleep 90; continue; fi
    break
done
echo "gpurun rc=$rc after $i attempt(s)" >> "$log"
tail -120 "$log"
