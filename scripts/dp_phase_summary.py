import re, sys, numpy as np
rows = []
for ln in open(sys.argv[1]):
    if not ln.startswith("dpprof"): continue
    ln = ln.replace("poll0", "pollz")
    rows.append([int(x) for x in re.findall(r"-?\d+", ln)])
names = ["stage","poll0","fold","prescan","waitleft","scan","finish","skipped"]
A = np.array(rows, dtype=np.float64)
tot = A[:,3].sum(); nr = A[:,12].sum()
print("alignments", len(A)//12, "cycles per row visit %.0f" % (tot/nr), "in-edges per row %.2f" % (A[:,13].sum()/nr))
for k,n in enumerate(names):
    print("%-9s %5.1f%%  %6.0f cycles/row" % (n, 100*A[:,4+k].sum()/tot, A[:,4+k].sum()/nr))
for w in range(12):
    B = A[A[:,2]==w]
    print("warp %2d rows %6.0f total %6.1fM  " % (w, B[:,12].mean(), B[:,3].mean()/1e6) + " ".join("%s %4.1f" % (n, 100*B[:,4+k].sum()/B[:,3].sum()) for k,n in enumerate(names)))
