echo "== K=63 base occ 13400"; python tools/profile_step.py --steps 3 --k 63 --bucket-occ 13400 2>&1 | tail -n 1 | cut -c1-900
echo "== K=63 base occ 9000"; python tools/profile_step.py --steps 3 --k 63 --bucket-occ 9000 2>&1 | tail -n 1 | cut -c1-900
echo "== K=31 base occ 36000 (2^15)"; python tools/profile_step.py --steps 3 --bucket-occ 36000 2>&1 | tail -n 1 | cut -c1-900
