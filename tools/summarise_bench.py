"""One summary line per bench JSON (a file holding the line itself, a pretty-printed copy, or a log whose last line is it)."""
import json, sys


def load(path):
    txt = open(path).read()
    try:
        return json.loads(txt)
    except Exception:
        for l in reversed(txt.splitlines()):
            l = l.strip()
            if l.startswith("{") and l.endswith("}"):
                try:
                    return json.loads(l)
                except Exception:
                    pass
    return None


for p in sys.argv[1:]:
    d = load(p)
    name = p.split("/")[-1]
    if not d or "value" not in d:
        print(name, "no bench line")
        continue
    r = d.get("roofline") or {}
    it = (r.get("iteration") or {})
    c3, c5 = d.get("c3_cg7_128") or {}, d.get("c5_bicgstab7cd_512") or {}
    par = d.get("parity") or {}
    ro = (par.get("reference_order") or {})
    e2e = d.get("e2e") or {}
    print(f"{name}: N={d.get('n_gpus')} {d['config'].get('workload')} it/s={d['value']:.1f} e2e={e2e.get('value', 0):.1f} spmv_ms={r.get('avg_launch_ms', 0):.4f} "
          f"spmv_frac={r.get('frac', 0):.3f} iter_frac={it.get('frac_of_peak', 0):.3f} c3={c3.get('value')} c5={c5.get('value')} "
          f"parity_rel={par.get('rel_l2')} ok={par.get('ok')} exact_bits={ro.get('bit_identical_solution')} clocks={(d.get('clocks') or {}).get('sm_mhz')} {(d.get('clocks') or {}).get('reasons')}")
