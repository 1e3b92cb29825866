for cls in Asian_GBM Digital_Merton; do
python tools/level_times.py --cls $cls --M 4 --Lmax 2 --n 1073741824
python tools/level_times.py --cls $cls --M 2 --Lmax 3 --n 1073741824
done
