bash -c 'sed -n "/cat > \/tmp\/san.py/,/^PY$/p" tools/runs/r2u_sanitizer.sh | sed "1d;\$d" > /tmp/san.py'
compute-sanitizer --tool initcheck --error-exitcode 7 --print-limit 10 python /tmp/san.py > gpurun_out/sanitizer_initcheck_r2.log 2>&1; echo "initcheck rc $?"
grep -E "Uninitialized|at |ERROR SUMMARY" gpurun_out/sanitizer_initcheck_r2.log | sort | uniq -c | sort -rn | head -20
