for gs in 4 6 8 12; do echo "== GRID_SMALL $gs"; PBRT_B200_GRID_SMALL=$gs python tools/step_diag.py 2>&1 | grep -E "plain"; done
for gs in 4 5 6 10 16; do echo "== GRID_SHADE $gs"; PBRT_B200_GRID_SHADE=$gs python tools/step_diag.py 2>&1 | grep -E "plain"; done
