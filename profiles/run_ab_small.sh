for cfg in "ISX_SMALL_MAXQ=4" "ISX_SMALL_MAXQ=8" "ISX_SMALL_PATH=0"; do
  echo "=== $cfg"
  env $cfg timeout 60 python profiles/prof_small.py --reps 10 2>&1 | grep "^1x\|mixed\|^8x\|host API" | cut -c1-150
done
echo "=== dbg"
ISX_SMALL_DEBUG=1 timeout 60 python profiles/dbg_small.py 2>&1 | grep "cta   0\|host api" | tail -8 | cut -c1-400
