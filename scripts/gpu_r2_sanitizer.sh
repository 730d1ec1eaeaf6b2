# round 2: compute-sanitizer over the smoke case (1 320 particles, 3 Fraser steps with thermal coupling, both arithmetic modes) and a contact + ghost + KickDrift case
tag=${1:-r2_sanitizer}
cat > /tmp/san_case.py <<PY
import sys
sys.path.insert(0, ".")
import __graft_entry__ as g
g.smoke()
from weldform_b200 import harness
from tests import cases
c = cases.perturbed(cases.compression(dx=0.03, nsteps=6, ghost=True), amp=60.0); c["solver"] = 1.0
harness.run_case(c, device=0)
c = cases.contact_compression(dx=0.0032, nsteps=60, vmax=5.0, tau=1e-5, dens=8, cfl=0.3)      # the contact_100 golden's case, 60 steps: friction stick + slide
harness.run_case(c, device=0)
c = cases.fsw_like(dx=0.0009, nsteps=20, L=0.02, rt=0.004)                                    # Johnson-Cook, thermal contact, rotating tool
harness.run_case(c, device=0)
print("sanitizer case done")
PY
for tool in memcheck racecheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_case.py > gpurun_out/sanitizer_${tool}_$tag.log 2>&1
  echo "== $tool: $(grep -c 'sanitizer case done' gpurun_out/sanitizer_${tool}_$tag.log) run(s) completed; $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/sanitizer_${tool}_$tag.log | tail -1)"
done
