for i in 1 2 3; do
  echo "== slots 4"; python tools/r2_c4_probe.py 2>&1 | grep "pinned " | awk '{printf "%s ", $2} END{print ""}'
  echo "== slots 2"; MAICOS_B200_PREF_SLOTS=2 python tools/r2_c4_probe.py 2>&1 | grep "pinned " | awk '{printf "%s ", $2} END{print ""}'
done
