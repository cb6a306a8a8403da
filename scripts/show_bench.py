import json, sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    print(f, "| asm %.0f Melem/s %.3f ms frac %.3f"%(d["value"], d["assembly"]["ms"], d["roofline"]["frac"]),
          "| pcg %.3g DOF-it/s %.3f ms/it frac %.3f"%(d["pcg"]["dof_iters_per_s"], d["pcg"]["ms_per_iter"], d["pcg"]["roofline"]["frac"]),
          "| e2e %.0f"%d["e2e"]["value"], "| plan_ms %.0f"%d["config"]["pattern_build_ms"], "| launches", d["gpu_launches"], "| n_gpus", d["n_gpus"])
    print("     clocks", d["clocks"], "\n     solve", d.get("solve"), "cpu", d.get("cpu_baseline",{}).get("value"))
