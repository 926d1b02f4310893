# patched reference (oracle/_ref/hsmc_gpu_patched) vs the drop-in driver: same input, compare every output file
d1=$(mktemp -d); d2=$(mktemp -d)
cat > $d1/in.dat <<'EOI'
rho 0.85
cells_x 16
cells_y 8
cells_z 8
type 2
neigh_list 1.05 10
dr_max 0.12
opt 1 60 6 0.5 0.5
press_virial 0.002 10
press_thermo 0.0001 0.002 10
rdf 0.02 4.0 20 100
widom 20000 10
ql 6 1.5 10
seed 31
restart_write 40
config_write 40 100
sweep_eq 40
sweep_stat 60
out 20
EOI
cp $d1/in.dat $d2/in.dat
(cd $d1 && /root/repo/oracle/_ref/hsmc_gpu_patched -i in.dat > out.txt 2>&1; echo "patched rc=$?"; tail -2 out.txt)
(cd $d2 && /root/repo/hsmc_b200/host/hsmc_b200 -i in.dat > out.txt 2>&1; echo "driver rc=$?"; tail -2 out.txt)
echo "files: $(ls $d1 | tr '\n' ' ') | $(ls $d2 | tr '\n' ' ')"
for f in $(ls $d1); do
  case $f in
    in.dat) ;;
    out.txt) diff <(grep -v "Elapsed time" $d1/$f) <(grep -v "Elapsed time" $d2/$f) > /dev/null && echo "SAME stdout" || { echo "DIFF stdout"; diff <(grep -v "Elapsed time" $d1/$f) <(grep -v "Elapsed time" $d2/$f) | head -10; } ;;
    *.gz) cmp <(zcat $d1/$f) <(zcat $d2/$f) > /dev/null && echo "SAME $f" || echo "DIFF $f" ;;
    restart_*) n=$(stat -c %s $d1/$f); cmp -n $n $d1/$f $d2/$f > /dev/null && echo "SAME $f (first $n bytes; driver file is $(stat -c %s $d2/$f))" || echo "DIFF $f" ;;
    *) cmp $d1/$f $d2/$f > /dev/null && echo "SAME $f" || { echo "DIFF $f"; diff $d1/$f $d2/$f | head -6; } ;;
  esac
done
