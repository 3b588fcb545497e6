# A/B harness: runs bench.py (kernel-only figure) for each library given, round-robin, REPS times
REPS=${REPS:-2}; ARGS=${ARGS:---no-cpu --pairs 500000 --steps 3}
for i in $(seq $REPS); do for lib in "$@"; do echo -n "$lib: "; KSW2B_LIB=$PWD/$lib python bench.py $ARGS 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"; done; done
