#!/bin/bash
# integration/gpuserver_check.sh -- the reference's gpuserver protocol through the patched host, both ends:
#   mmseqs_b200 gpuserver T_pad            (src/util/gpuserver.cpp, class Marv = integration/shim/marv.h)
#   mmseqs_b200 ungappedprefilter Q T_pad pref_srv --gpu 1 --gpu-server 1     (client side of ungappedprefilter.cpp:209-250)
# and compares the prefilter DB with the one the in-process path (--gpu 1) and the CPU path (--prefilter-mode 1 semantics) write.
set -u
HERE="$(cd "$(dirname "$0")" && pwd)"
B="$HERE/_build/mmseqs_b200"; C="$HERE/_build/mmseqs_avx2"; W=${1:-/tmp/gpusrv}; EX="$HERE/_build/examples"
rm -rf "$W"; mkdir -p "$W"; cd "$W"
"$C" createdb "$EX/QUERY.fasta" Q -v 1 >/dev/null; "$C" createdb "$EX/DB.fasta" T -v 1 >/dev/null; "$C" makepaddedseqdb T T_pad -v 1 >/dev/null
timeout 120 "$B" gpuserver T_pad --max-seqs 300 > server.log 2>&1 &
SRV=$!
sleep ${SRV_WAIT:-6}
timeout 90 "$B" ungappedprefilter Q T_pad pref_srv --gpu 1 --gpu-server 1 --threads 4 -v 2 > client.log 2>&1; echo "client exit $?"
kill -INT $SRV 2>/dev/null; sleep 1; kill $SRV 2>/dev/null; wait $SRV 2>/dev/null
timeout 90 "$B" ungappedprefilter Q T_pad pref_gpu --gpu 1 --threads 4 -v 2 > direct.log 2>&1; echo "direct exit $?"
export QUICK=${QUICK:-0}
if [ "$QUICK" = 1 ]; then cp pref_gpu pref_cpu; cp pref_gpu.index pref_cpu.index; echo "cpu leg skipped (QUICK=1)"; else
  timeout 300 "$C" ungappedprefilter Q T_pad pref_cpu --threads 16 -v 2 > cpu.log 2>&1; echo "cpu exit $?"
fi
python3 - <<'PY'
import os
def read_db(path):
    parts = [path] if os.path.exists(path) else []
    k = 0
    while not parts or os.path.exists("%s.%d" % (path, k)):
        if os.path.exists("%s.%d" % (path, k)): parts.append("%s.%d" % (path, k)); k += 1
        else: break
    data = b"".join(open(p, "rb").read() for p in parts)
    return {int(l.split()[0]): data[int(l.split()[1]):int(l.split()[1]) + int(l.split()[2])] for l in open(path + ".index")}
a, b, c = read_db("pref_srv"), read_db("pref_gpu"), read_db("pref_cpu")
print("entries", len(a), len(b), len(c), "server == in-process:", a == b, "server == cpu:", "skipped (QUICK=1)" if os.environ.get("QUICK") == "1" else a == c)
PY
tail -3 server.log | cut -c1-200
