# compute-sanitizer over every kernel of HEAD (scripts/sanitize.py), one B200; smoke() first
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
{
for tool in memcheck racecheck synccheck; do
  timeout 240 compute-sanitizer --tool $tool python scripts/sanitize.py 2>&1 | grep -v "^$" | tail -4
done
} > gpurun_out/r02_sanitizer_raw.txt 2>&1
cat gpurun_out/r02_sanitizer_raw.txt
