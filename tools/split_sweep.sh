for T in 4736 9472 18944 37888; do for C in 32 64 128 256; do
echo "== target $T chain $C"; NDP_SPLIT=$T NDP_SPLIT_CHAIN=$C timeout 100 python tools/perf_probe.py rsaFACT-32bit:d:12:e rsaFACT-32bit:d:12 rsaFACT-32bit:q:1024 2>&1 | grep rep1 | sed -E 's/.*(rsaFACT[^ ]*) rep1.*DFS wall ([0-9.]+)s kernel ([0-9.]+)ms.*/\1 dfs_kernel_ms \3/'
done; done
