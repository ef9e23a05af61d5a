echo "== round 1"; PYTHONPATH=$PWD/r1ref python r1ref/tools/rep_cases.py 2>&1 | head -5
echo "== current"; python tools/rep_cases.py 2>&1 | head -5
for v in lag2 lag6 ch1 ch4 noabs; do echo "== $v"; PQB200_LIB=$PWD/piquassoboost_b200/libpqb200_$v.so python tools/rep_cases.py 2>&1 | head -2; done
echo "== round 1 again"; PYTHONPATH=$PWD/r1ref python r1ref/tools/rep_cases.py 2>&1 | head -2
