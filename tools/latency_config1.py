import sys, time, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from olap_in_memory_b200 import Cube, GenericDimension, TimeDimension, _native as N
from olap_in_memory_b200.store import GpuStore
N.init(0)
cube = Cube([TimeDimension('time','day','2010-01-01','2019-12-31'), GenericDimension('location','city',['paris','tokyo'])])
cube.createStoredMeasure('m1', {'time':'sum'}, 'float32', 0)
cube.setData('m1', np.random.default_rng(0).integers(1,100,cube.storeSize).astype(np.float32))
for name, fn in (('Cube.drillUp day->month (7304 cells)', lambda: cube.drillUp('time','month')),
                 ('Cube.dice', lambda: cube.dice('location','city',['tokyo'])),
                 ('Cube.reorderDimensions', lambda: cube.reorderDimensions(['location','time'])),
                 ('getData', lambda: cube.getData('m1')),
                 ('store.total', lambda: cube.getTotal('m1'))):
    fn(); 
    t=time.perf_counter()
    for _ in range(200): fn()
    dt=(time.perf_counter()-t)/200*1e6
    print(f'{name:45s} {dt:8.1f} us/call  kernel {N.lib().olap_last_op_ms()*1e3:6.1f} us path {N.lib().olap_last_op_path().decode()}')
s = cube.storedMeasures['m1']
day = cube.dimensions[0]; m = day.getGroupIndexFromRootIndexMap('month'); ident=np.arange(2,dtype=np.int32)
t=time.perf_counter()
for _ in range(500): GpuStore.drillUp_lowered([s],[3652,2],[120,2],[m,ident],['sum'])
print('lowered drillUp (C ABI + ctypes)', (time.perf_counter()-t)/500*1e6, 'us/call')
