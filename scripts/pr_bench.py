import json,sys
b=json.loads(sys.stdin.read())
r=b['roofline']
print('value %.0f wall %.2f dev %.2f slicer_stage %.2f kernel/step %.2f launches/step %.1f frac %.3f' % (b['value'], b['ms_per_step'], b['device_ms_per_step'], b['slicer_ms_per_step'], r['avg_launch_ms']*r['launches_per_step'], r['launches_per_step'], r['frac']))
