echo "== K=63"; python tools/profile_step.py --steps 3 --k 63 2>&1 | tail -n 1 | cut -c1-900
echo "== K=47"; python tools/profile_step.py --steps 3 --k 47 2>&1 | tail -n 1 | cut -c1-900
python -m pytest tests -m gpu -x -q -k "not full_size and not fullsize and not multi" 2>&1 | tail -n 2
