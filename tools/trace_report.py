"""Pretty-prints the clock64 timeline written by a -DFA_TRACE build ($FA_B200_TRACE)."""
import sys
names = ["s0_ready", "s0_ld", "s0_pA", "s0_pB", "s1_ready", "s1_ld", "s1_pA", "s1_pB",
         "pv0A", "pv0B", "qk0", "pv1A", "pv1B", "qk1", "s0_max", "s1_max"]
rows = [[int(x) for x in l.split()] for l in open(sys.argv[1]) if l.strip() and not l.startswith(("I ", "#", "B2"))]
t0 = min(v for r in rows for v in r if v > 0)
print("j    " + " ".join(f"{n:>8s}" for n in names))
for j, r in enumerate(rows[:int(sys.argv[2]) if len(sys.argv) > 2 else 24]):
    print(f"{j:<4d} " + " ".join(f"{(v - t0) if v else -1:8d}" for v in r[:16]))
# per-iteration deltas in steady state
print("\nsteady-state deltas (iterations 8..20):")
for a, b in [("s0_ready", "s0_ld"), ("s0_ld", "s0_max"), ("s0_max", "s0_pA"), ("s1_ld", "s1_max"), ("s1_max", "s1_pA"), ("s1_pA", "s1_pB"), ("s0_pA", "s0_pB"), ("s0_pB", "pv0B"), ("pv0A", "qk0"),
             ("s0_ready", "s1_ready"), ("s1_ready", "s1_pB"), ("qk0", "pv1A"), ("pv1A", "qk1")]:
    ia, ib = names.index(a), names.index(b)
    d = [rows[j][ib] - rows[j][ia] for j in range(8, 20) if rows[j][ia] and rows[j][ib]]
    if d: print(f"  {a:>9s} -> {b:<9s} mean {sum(d)/len(d):8.0f}  min {min(d)} max {max(d)}")
per = [rows[j + 1][0] - rows[j][0] for j in range(8, 20) if rows[j][0] and rows[j + 1][0]]
print("  period (s0_ready j -> j+1): mean", sum(per) / len(per))
