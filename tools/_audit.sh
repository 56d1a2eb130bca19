cd $GRAFT_REPO_ROOT
M=smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dsetp_pred_on.sum
for c in $(python tools/audit_fp64.py list); do
  python tools/audit_fp64.py run $c > gpurun_out/audit_$c.txt 2> gpurun_out/audit_$c.err && \
  ncu --metrics $M --clock-control none --csv --log-file gpurun_out/audit_$c.csv python tools/audit_fp64.py run $c > /dev/null 2>&1
  tail -1 gpurun_out/audit_$c.txt
done
python tools/audit_fp64.py table gpurun_out
