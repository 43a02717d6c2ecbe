"""Print the launch ID of the longest kernel whose name contains argv[2] in an ncu launch csv."""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = rows[0]
best, bid = -1.0, 0
for r in rows[1:]:
    d = dict(zip(h, r))
    if sys.argv[2] in d["Kernel Name"]:
        v = float(d["Metric Value"].replace(",", ""))
        if v > best:
            best, bid = v, int(d["ID"])
print(bid)
