cd $GRAFT_REPO_ROOT
for e in "1 0" "2 0" "2 1" "3 0" "5 0"; do
  timeout 60 ./profiles/micro/tc_probe $e 2>&1 | tail -12
  echo "rc=$?"
done
