import json, sys
for l in open(sys.argv[1]):
    l = l.strip()
    if not l.startswith('{'):
        print(l); continue
    d = json.loads(l)
    keys = ('case', 'rel_err', 'nan', 'timeout_code', 'ms', 'error', 'bad_count', 'dense_tflops_x3', 'same_nan_pattern', 'inf', 'max_diff_finite')
    print({k: (round(v, 4) if k in ('ms', 'dense_tflops_x3') else v) for k, v in d.items() if k in keys and not (k == 'timeout_code' and v == 0) and not (k == 'nan' and v == 0)})
