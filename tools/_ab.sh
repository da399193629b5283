set -e
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests -m gpu -q -x -k "gather or rotate or chain or config3 or pipeline" 2>&1 | tail -2
timeout 200 python tools/bench_ops.py --quick 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('B(256,6):', d['config3']['fused']['images/s'], d['config4'])"
