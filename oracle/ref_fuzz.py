"""One randomised job, oracle against the reference's own code (oracle/_ref), bit for bit:
    for s in $(seq 100 139); do timeout -s KILL 25 python oracle/ref_fuzz.py $s | tail -1; done
TEST INFRASTRUCTURE ONLY.  Run each job under `timeout`: the reference's pairwise exchange can
deadlock under the shim when ranks need different mailbox sizes.  The seed picks
np in {1,2,4,8}, the method, unique, label density, zeros and flags; 37 cases per job cover
plain / vec / many, six dom x op pairs (incl. gs_bpr) and both transposes.  Round 1: seeds 100-139 gave 38
identical jobs, one reference failure (126) and one reference hang (131)."""
import sys, itertools
import os; _R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0]=[os.path.join(_R,'exaflow-exags-upc_b200'),os.path.join(_R,'oracle'),os.path.join(_R,'tests')]
import numpy as np, oracle as O, ref, cases
DT=[np.float64,np.float32,np.int32,np.int64]
seed=int(sys.argv[1]); bad=0
rng=np.random.default_rng(seed)
np_=int(rng.choice([1,2,4,8])); unique=int(rng.random()<0.3); method=int(rng.choice([1,1,3]))
nl=int(rng.integers(3,200))
ids=[cases.random_ids(int(rng.integers(0,400))+5, nl, 1000*seed+r, p_zero=float(rng.choice([0,0.2])), p_flag=0.0 if unique else float(rng.choice([0,0.25,0.8]))) for r in range(np_)]
combos=[(1,8,0,0,0)]+[(m,vn,d,o,t) for (m,vn),(d,o),t in itertools.product(((0,1),(1,3),(2,2)),((0,0),(2,3),(3,2),(1,1),(0,2),(2,4)),(0,1))]      # the last pair: gs_bpr on int
data=[[cases.values(ids[r].size*(1 if m==0 else vn), DT[d], cases.OPS_BPR[o], 31*k+r) for r in range(np_)] for k,(m,vn,d,o,t) in enumerate(combos)]
mine=[[a.copy() for a in row] for row in data]
maps,_=ref.run(ids,combos,data,unique=unique,method=method,want_maps=True)
orc=O.Oracle(ids,unique=unique,method=O.METHOD_ALLREDUCE if method==3 else O.METHOD_PAIRWISE)
for r in range(np_):
    for w in range(3):
        if not np.array_equal(orc.map(r,w),maps[r][w]): bad+=1; print("map mismatch",seed,r,w)
for k,(m,vn,d,o,t) in enumerate(combos):
    us=[[row[i*ids[r].size:(i+1)*ids[r].size] for i in range(vn)] if m==2 else row for r,row in enumerate(mine[k])]
    orc.exec(us,m,vn,d,o,t)
    for r in range(np_):
        if DT[d] in (np.float64,np.float32) and o in (0,1) and method==3:
            try: np.testing.assert_array_max_ulp(mine[k][r],data[k][r],maxulp=4)
            except AssertionError: bad+=1; print("ulp mismatch",seed,k,r)
        elif not np.array_equal(mine[k][r],data[k][r]): bad+=1; print("value mismatch",seed,k,r,combos[k])
print(f"seed {seed} np={np_} unique={unique} method={method}: bad={bad}")
sys.exit(1 if bad else 0)
