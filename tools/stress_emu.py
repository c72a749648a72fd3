"""Randomized stress of the engine logic on the CPU: tests/emu (the device functions compiled with g++) against the
sequential oracle over random PHOLD / MachRep2 / MachRep1 configurations (window, batch, ring size, skew, partitions,
lagging GVT).  Usage: python tools/stress_emu.py <seed> <seconds> [fast] [simt].  `simt` runs the warp-level build (32 host
threads as the lanes of a warp: work queue, deferred delivery, whole-warp search; random tiles per warp).  Ring overflows are legitimate outcomes;
`fast` keeps every case small (used by tests/test_stress_short.py).  Exit status 1 when a case fails."""
import sys, time
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import helpers as H, numpy as np
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 1)
t0=time.time(); n_ok=0; n_over=0; fails=[]
budget=float(sys.argv[2]) if len(sys.argv)>2 else 300
fast = 'fast' in sys.argv[3:]
simt = 'simt' in sys.argv[3:]
def run_emu(p, sd, every):
    if simt:
        return H.emu_simt_run(p, sd, every, int(rng.choice([1, 1, 2, 3])))
    return H.emu_run(p, sd, gvt_every=every)
while time.time()-t0 < budget:
    kind = rng.integers(0,10)
    if kind < 7:
        n = int(rng.integers(1, 300 if fast else 1500)); pop = int(rng.integers(1, 12)); rp = int(rng.choice([0, 5, 30, 50, 100]))
        hz = float(rng.choice([0.0, 0.05, 0.1, 0.3, 0.8, 2.5, -1.0])); batch = int(rng.choice([1,2,3,4,5,6,7,8,12,16,33,64]))
        n_hot = int(rng.choice([0,0,1,3,16])); hot = int(rng.choice([5,30,80])) if n_hot else 0
        la = float(rng.choice([0.0, 0.1, 0.5])); mean = float(rng.choice([0.0, 0.2, 1.0, 3.0]))
        if la == 0.0 and mean == 0.0: mean = 1.0
        cap = int(rng.choice([64,128,512,2048])); world = int(rng.choice([1,1,2,3,8]))
        pl = int(rng.choice([0,0,4,16])); fo = int(rng.choice([0,0,2,8]))
        seed = int(rng.integers(0, 2**31)); T = float(rng.choice([1.0, 2.0] if fast else [1.0, 3.0, 6.0]))
        if fast and (hz == 0.0 or la == 0.0): hz, la = 0.3, 0.1          # zero windows / lookaheads take 10^5 tiny steps
        every = int(rng.choice([1,1,1,3,11]))
        p = H.phold_params(n, T, population=pop, remote_pct=rp, lookahead=la, mean=mean, seed=seed, n_hot=min(n_hot,n), hot_pct=hot,
                           horizon=hz, cap=cap, batch=batch, world=world, proc_limit=pl, fossil=fo)
        desc = dict(n=n,pop=pop,rp=rp,hz=hz,batch=batch,n_hot=n_hot,hot=hot,la=la,mean=mean,cap=cap,world=world,pl=pl,fo=fo,seed=seed,T=T,every=every)
        print("case", desc, flush=True)
        o_exec,o_hash,o_sum = H.oracle_phold(p)
        rc,e,eh,_,st = run_emu(p, int(rng.integers(0,1000)), every)
        if rc == 5: n_over += 1; continue
        ok = rc==0 and (e==o_exec).all() and (eh==o_hash).all() and st.executed-st.rolled_back==st.committed
    else:
        nrep = int(rng.integers(1, 12)); hz = float(rng.choice([0.0, 0.5, 2.0, 50.0, -1.0])); batch = int(rng.choice([1,2,4,6,8,16]))
        cap = int(rng.choice([64,256,1024])); world = int(rng.choice([1,2,3])); T = float(rng.choice([20.0, 40.0] if fast else [30.0, 120.0]))
        if fast and hz == 0.0: hz = 0.5
        m1 = bool(rng.integers(0,2)); every = int(rng.choice([1,1,5]))
        s1 = int(rng.integers(1, 10**6)); s2 = int(rng.integers(1, 10**6))
        desc = dict(model='m1' if m1 else 'm2', nrep=nrep,hz=hz,batch=batch,cap=cap,world=world,T=T,s1=s1,s2=s2,every=every)
        if m1:
            p = H.machrep1_params(nrep, T, horizon=hz, cap=cap, batch=batch, world=world, seed1=s1, seed2=s2); res = H.oracle_machrep1(p)
        else:
            p = H.machrep_params(nrep, T, horizon=hz, cap=cap, batch=batch, world=world, seed1=s1, seed2=s2); res = H.oracle_machrep(p)
        rc,ne,h,mr,st = run_emu(p, int(rng.integers(0,1000)), every)
        if rc == 5: n_over += 1; continue
        ok = rc==0 and all(res[r].up_time_prop==mr[r].up_time_prop and res[r].total_up_time==mr[r].total_up_time and
                           [res[r].hash[k] for k in range(3)]==[int(h[3*r+k]) for k in range(3)] for r in range(nrep))
    if ok: n_ok += 1
    else:
        fails.append((rc, desc)); print('FAIL', rc, desc, flush=True)
        if len(fails) > 5: break
print('ok', n_ok, 'overflow (legit)', n_over, 'fails', len(fails), 'in', round(time.time()-t0), 's')
sys.exit(1 if fails else 0)
