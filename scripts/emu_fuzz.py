"""Randomised sweep of dq_cuda_bsdiff_streams on the CPU logic emulator (tests/emu: the CUDA sources compiled for the host;
test infrastructure) against the oracle: mid-size pairs with zero runs, periodic records, moved / inserted / deleted
blocks and point damage, a random host-thread shape per pair, certified stretches compared byte for byte.

    python scripts/emu_fuzz.py SEED SECONDS        # prints "seed S: N pairs, 0 mismatches"; a failing pair is saved to /tmp
"""
import sys, time, os, numpy as np
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
os.environ["DQ_CHECK_CERTS"]="1"
def pair(rng):
    n = int(rng.integers(20_000, 500_000))
    sigma = int(rng.choice([2, 4, 16, 256]))
    old = rng.integers(0, sigma, n, dtype=np.uint8)
    a = int(rng.integers(0, n // 2)); old[a:a + int(rng.integers(100, n//4))] = 0
    b = int(rng.integers(0, n // 2)); rec = rng.integers(0, 256, int(rng.integers(1, 64)), dtype=np.uint8)
    k = int(rng.integers(100, n//4)); old[b:b + k] = np.resize(rec, k)[:old[b:b + k].size]
    parts, cur = [], 0
    for c in np.sort(rng.integers(0, n, int(rng.integers(1, 40)))):
        c = int(max(c, cur)); parts.append(old[cur:c]); op = int(rng.integers(0, 4)); k = int(rng.integers(1, 20_000))
        if op == 0: parts.append(rng.integers(0, sigma, k, dtype=np.uint8)); cur = min(n, c + k)
        elif op == 1: parts.append(rng.integers(0, 256, k, dtype=np.uint8)); cur = c
        elif op == 2: cur = min(n, c + k)
        else:
            s = int(rng.integers(0, max(1, n - k))); parts.append(old[s:s + k]); cur = c
    parts.append(old[cur:])
    new = np.concatenate(parts).copy()
    if new.size:
        hits = rng.integers(0, new.size, int(rng.integers(0, 200))); new[hits] ^= 1
    return old, new
if __name__ == '__main__':
    import emu, oracle
    from deltaq_b200 import CudaSuffixSort, bsdiff
    seed=int(sys.argv[1]); budget=float(sys.argv[2])
    rng=np.random.default_rng(seed)
    t0=time.time(); cnt=0; bad=0
    while time.time()-t0 < budget:
        shape=rng.choice(["0,1","1,1","3,2","7,3","2,1,4,8"])
        os.environ["DQ_HOST_THREADS"]=str(shape)
        s=CudaSuffixSort(_lib=emu.library())
        old,new=pair(rng)
        got=bsdiff.create_streams(old,new,s)
        ref=oracle.bsdiff_streams(old,new)
        ok=all(got[k]==ref[k] for k in ("ctrl","diff","extra")) and got["search_visits"]==ref["search_calls"]
        if not ok:
            bad+=1; np.savez(f"/tmp/emu_fuzz_bad_{seed}_{cnt}.npz",old=old,new=new); print("MISMATCH",seed,cnt,shape,flush=True)
        s.dispose(); cnt+=1
    print(f"seed {seed}: {cnt} pairs, {bad} mismatches",flush=True)
