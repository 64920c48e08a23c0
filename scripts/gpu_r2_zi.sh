ETR_DENSE_IEEE=1 timeout 120 python scripts/dbg_extras.py keras denorm 2>&1 | tail -1 | sed 's/^/IEEE, subnormal m: /'
timeout 120 python scripts/dbg_extras.py keras denorm 2>&1 | tail -1 | sed 's/^/MUFU, subnormal m: /'
timeout 200 python -m pytest tests -m gpu -q -k "adam or keras or checkpoint or oracle" 2>&1 | tail -2
