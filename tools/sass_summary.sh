#!/bin/bash
# SASS evidence for the copy-engine / async-copy / peer-atomic paths: per kernel of libfinrl_b200.so, how many
# UBLKCP (cp.async.bulk = TMA), SYNCS (mbarrier), LDGSTS (cp.async), RED/ATOM (statistics, peer push) instructions.
lib=${1:-finrl_b200/libfinrl_b200.so}
echo "# $(basename $lib): $(cuobjdump -lelf $lib | grep -c cubin) cubin(s), archs: $(cuobjdump -lelf $lib | grep -o 'sm_[0-9a]*' | sort -u | tr '\n' ' ')"
cuobjdump -sass $lib | awk '
/Function :/ { fn=$3; next }
/UBLKCP\.S\.G/ { ld[fn]++ } /UBLKCP\.G\.S/ { st[fn]++ } /SYNCS/ { sy[fn]++ } /LDGSTS/ { cp[fn]++ }
/ RED\.| REDG\.|ATOMG\.|ATOM\./ { at[fn]++ } /\.SYS/ { sys[fn]++ } /FENCE\.VIEW\.ASYNC|FENCE.*PROXY/ { fe[fn]++ }
END { for (f in ld) k[f]=1; for (f in st) k[f]=1; for (f in cp) k[f]=1; for (f in sys) k[f]=1;
      printf "%-8s %-8s %-6s %-7s %-9s %-5s %s\n", "UBLKCP.L", "UBLKCP.S", "SYNCS", "LDGSTS", "RED/ATOM", ".SYS", "kernel";
      for (f in k) printf "%-8d %-8d %-6d %-7d %-9d %-5d %s\n", ld[f], st[f], sy[f], cp[f], at[f], sys[f], f }' | (read h; echo "$h"; sort -k7 | c++filt | cut -c1-170)
