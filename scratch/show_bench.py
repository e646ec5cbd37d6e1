import json, sys
d = json.load(open(sys.argv[1]))
print('value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'cpu', d.get('cpu_baseline', {}).get('value'))
print(json.dumps(d.get('other_configs'), indent=1))
