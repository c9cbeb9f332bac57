import sys,json
for line in sys.stdin:
    line=line.strip()
    if not line.startswith("{"): print(line); continue
    d=json.loads(line); print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k in ("kind","M","spmm_ms","spmm_frac","taylor_ms","taylor_frac","taylor_steps_s","auto_ms","auto_K","auto_frac","auto_steps_s","cheb_K","cheb_steps_s","taylor_K")})
