"""BASELINE.json configs[3]: equation-of-state scan with the volumes distributed one System per GPU (torchrun) --
fcc Al, WT + PBE, 2000 eV cutoff (docs/source/example_elastic.rst:81-86: V0 = 16.76389 A^3, E0 = -57.18370 eV,
K0 = 78.80961 GPa for the primitive cell)."""
import json, os, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import profess_ad_b200.functionals as F
from profess_ad_b200 import parallel
from profess_ad_b200.crystal_tools import get_cell
from profess_ad_b200.system import System

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    os.environ.setdefault('NCCL_DEBUG_FILE', '/tmp/nccl.%h.%p.log')
    dist.init_process_group('nccl', device_id=dev)
pot = os.path.join(ROOT, 'tests', 'potentials', 'al.gga.recpot')


def make_system():
    box, frac = get_cell('fcc', vol_per_atom=16.9, coord_type='fractional')
    terms = [F.IonIon, F.IonElectron, F.Hartree, F.WangTeter, F.PerdewBurkeErnzerhof]
    s = System(box, System.ecut2shape(2000, box), [['Al', pot, frac]], terms, units='a', coord_type='fractional', device=dev)
    s.optimize_density(ntol=1e-10)
    return s


torch.cuda.synchronize()
t0 = time.perf_counter()
params, err = parallel.eos_fit(make_system, f=0.05, N=max(9, world), eos='bm')
torch.cuda.synchronize()
if rank == 0:
    print(json.dumps({'world': world, 'seconds': time.perf_counter() - t0, 'K0_GPa': params[0], 'K0p': params[1], 'E0_eV': params[2],
                      'V0_A3': params[3], 'docs': {'K0_GPa': 78.80961, 'E0_eV': -57.18370, 'V0_A3': 16.76389}}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
