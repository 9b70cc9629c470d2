import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable:", e); continue
    t = d.get("timing", {})
    print(f"{f}: value {d['value']:.0f} ({d['ms_per_step']:.3f} ms)  e2e {d['e2e']['value']:.0f} ({d['e2e'].get('ms_per_step', 0):.3f} ms)  "
          f"e2e_pipe {d.get('e2e_pipelined', {}).get('value', 0):.0f}  lib_total {t.get('lib_total_ms_mean', 0):.3f} screen {t.get('lib_screen_ms_mean', 0):.3f}  "
          f"fb {d['config'].get('fallback_queries_in_timed_region')}  parity {d.get('parity')}  clocks {d.get('clocks', {}).get('sm_mhz')}")
