// Times the REAL reference package (Growblocks/olap-in-memory, single-threaded Node) on the
// operations of the hot path, with the same counter-based synthetic data as bench.py
// (SURVEY.md §8d).  UNEXECUTED in this repository's build image: there is no Node.js there
// (the CPU figures in profiles/ come from the C port of in-memory.js, labelled "port").
//
//   cd <checkout of Growblocks/olap-in-memory> && npm ci
//   node <this repo>/bench/node/bench_reference.js [cells=1000000] [repeats=5]
//
// Sizes: the store is a JS Map, which V8 caps at 2^24 entries (SURVEY.md F5), so
// cells <= 16 000 000; config 1 of BASELINE.json is `cells=7304`.
// Prints one JSON line per operation: {"op", "cells_in", "measures", "ms", "cells_per_s"}.
const path = require('path');
const { Cube, GenericDimension, TimeDimension } = require(path.resolve(process.cwd(), 'src'));

const cells = Number(process.argv[2] || 1e6);
const repeats = Number(process.argv[3] || 5);

// splitmix64 on BigInt, as bench.py: value = float32(1 + (h >> 40) * 999 / 2^24), never 0 / NaN
const MASK = (1n << 64n) - 1n;
function splitmix64(x) {
  x = (x + 0x9e3779b97f4a7c15n) & MASK;
  x = ((x ^ (x >> 30n)) * 0xbf58476d1ce4e5b9n) & MASK;
  x = ((x ^ (x >> 27n)) * 0x94d049bb133111ebn) & MASK;
  return x ^ (x >> 31n);
}
function synth(n, seed, measure) {
  const out = new Array(n);
  for (let i = 0; i < n; ++i) {
    const h = splitmix64(BigInt(i) ^ BigInt(seed) ^ (BigInt(measure) << 40n));
    out[i] = Math.fround(1 + Number(h >> 40n) * (999 / 16777216));
  }
  return out;
}

// [time day 2010-01-01..2019-12-31 (3652), g (cells / 3652 items)] : config 1 / config 2 shape
const inner = Math.max(1, Math.round(cells / 3652));
const items = Array.from({ length: inner }, (_, i) => `i${i}`);
const group = new GenericDimension('g', 'item', items);
group.addAttribute('item', 'parity', (item) => (Number(item.slice(1)) % 2 ? 'odd' : 'even'));
const time = new TimeDimension('time', 'day', '2010-01-01', '2019-12-31');
const cube = new Cube([time, group]);
const methods = ['sum', 'average', 'highest'];
methods.forEach((method, m) => {
  cube.createStoredMeasure(`m_${method}`, { time: method, g: method }, 'float32', 0);
  cube.setData(`m_${method}`, synth(cube.storeSize, 1, m));
});
cube.createComputedMeasure('ratio', '(m_sum + m_average) / m_highest');

function time_op(op, fn, cellsIn, measures) {
  fn(); // warm-up (JIT)
  const t0 = process.hrtime.bigint();
  for (let r = 0; r < repeats; ++r) fn();
  const ms = Number(process.hrtime.bigint() - t0) / 1e6 / repeats;
  console.log(JSON.stringify({ op, cells_in: cellsIn, measures, ms, cells_per_s: (cellsIn * measures) / (ms / 1e3) }));
}

const n = cube.storeSize;
const months = cube.drillUp('time', 'month');
time_op('drillUp time day->month', () => cube.drillUp('time', 'month'), n, 3);
time_op('drillUp time day->all', () => cube.drillUp('time', 'all'), n, 3);
time_op('drillUp g item->parity', () => cube.drillUp('g', 'parity'), n, 3);
time_op('dice g every other item', () => cube.dice('g', 'item', items.filter((_, i) => i % 2 === 0)), n, 3);
time_op('diceRange time one year', () => cube.diceRange('time', 'day', '2012-01-01', '2012-12-31'), n, 3);
time_op('reorderDimensions [g, time]', () => cube.reorderDimensions(['g', 'time']), n, 3);
time_op('drillDown time month->day', () => months.drillDown('time', 'day'), months.storeSize, 3);
time_op('getData stored', () => cube.getData('m_sum'), n, 1);
time_op('getData computed', () => cube.getData('ratio'), n, 1);
time_op('collapse', () => cube.collapse(), n, 3);
