# A/B on one box: last-warp refill (PQG_TILE_SYNC=0) vs CTA-wide barrier per tile (=1); unset = per-kernel default
for v in 0 1 0 1; do
  echo "== PQG_TILE_SYNC=$v"
  PQG_TILE_SYNC=$v python scripts/regex_probe.py 20000000 2>&1 | grep states | head -1
  PQG_TILE_SYNC=$v python scripts/bench_optional.py 20000000 2>/dev/null | python -c "
import json,sys
for r in json.load(sys.stdin)['results']: print(' ', r['column'], round(r['ms'],4), round(r['tiles_ms'],4))"
done
