"""Per-kernel SASS evidence for the tensor-core / TMA paths: counts of UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st),
UTMALDG / UTMASTG (TMA tensor copies), UBLKCP (bulk copies), SYNCS (mbarrier) in snd-vae_b200/libsndvae.so.
usage: python tools/sass_summary.py > profiles/sass_r2.txt   (runs cuobjdump -sass; no GPU needed)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "snd-vae_b200", "libsndvae.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = {"UTC*MMA": r"\bUTC[A-Z]*MMA", "LDTM": r"\bLDTM", "STTM": r"\bSTTM", "UTMALDG": r"\bUTMALDG", "UTMASTG": r"\bUTMASTG",
       "UBLKCP": r"\bUBLKCP", "SYNCS": r"\bSYNCS", "LDGSTS": r"\bLDGSTS"}
cnt = collections.OrderedDict(); cur = None; arch = set()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        cnt[cur] = collections.Counter(); continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m: arch.add(m.group(1))
    if cur:
        for k, p in pat.items():
            if re.search(p, line): cnt[cur][k] += 1
sha = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
print(f"# cuobjdump -sass snd-vae_b200/libsndvae.so (built from {sha}); images: {sorted(arch)}; kernels: {len(cnt)}")
print("# kernel | " + " | ".join(pat))
tot = collections.Counter()
for k, c in cnt.items():
    tot.update(c)
    if any(c[x] for x in ("UTC*MMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP")):
        print(f"{k[:90]} | " + " | ".join(str(c[x]) for x in pat))
print("# total | " + " | ".join(str(tot[x]) for x in pat))
