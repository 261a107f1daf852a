# A/B of the beam-grid cell size (development aid): LRT_OPT_BEAM_CELL_PCT = option 9
for o in "" "9=35" "9=50" "9=70" "1=3"; do echo "OPTIONS=$o"; LIDAR_RT_B200_OPTIONS="$o" timeout 300 python scripts/profile_frame.py --frames 6; done
