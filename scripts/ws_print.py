import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
for r in rows[1:]:
    print("  ", r[h.index("ID")], r[h.index("Metric Name")], r[h.index("Metric Value")])
