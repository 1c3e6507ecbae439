# usage: bash scripts/gemm_try.sh "M,N,K,beta,BN,splits" ...   (in-graph kernel spans of one shape under a forced tile / split)
for cfg in "$@"; do
  MRN_GEMM_TRY=$cfg MRN_GEMM_PROFILE_DUMP=gpurun_out/try.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/try.json 2>/dev/null
  python - "$cfg" <<'PY'
import csv,sys,json,collections
cfg=sys.argv[1].split(',')
M,N,K,beta=cfg[0],cfg[1],cfg[2],float(cfg[3])
for f in ('gpurun_out/try.csv.spans','gpurun_out/try.csv'):
    agg=collections.defaultdict(lambda:[0,0.0])
    try:
        for r in csv.reader(open(f)):
            if r[0]==M and r[1]==N and r[2]==K and abs(float(r[7])-beta)<1e-6:
                a=agg[(r[4],r[5],r[6])]; a[0]+=1; a[1]+=float(r[-1])
    except Exception as e:
        print(f, e); continue
    print(f.split('.')[-1], sys.argv[1], {k:(v[0], round(v[1]/v[0],1)) for k,v in agg.items()})
d=json.loads(open('gpurun_out/try.json').read().strip().splitlines()[-1])
print("   step ms", round(d['ms_per_step'],3))
PY
done
