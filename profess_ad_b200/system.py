"""``System``: the facade of PROFESS-AD (src/professad/system.py) on the B200-native hot path.

Same constructor, setters/getters and ``optimize_density`` keyword arguments as the reference, so
user scripts switch by changing the import.  What differs is underneath:

* every native energy term is one C-ABI call that returns the energy and the analytic potential;
* when all terms are native, ``optimize_density`` / ``functional_derivative`` use the *fused
  evaluator* (one pass structure for the whole term list, shared FFTs, chi-projection kernel) and
  the device-resident L-BFGS / TPGD loop (``_density_opt.py``) -- no host synchronisation inside an
  outer iteration;
* user-defined Python terms (lambdas, ``nn.Module.forward``) still work through the generic
  autograd closure, exactly as in the reference (system.py:830-854).

First derivatives are analytic as well: the ionic potential, the
ion-electron forces and the stress are native reductions (``csrc/ions.cu``, ``csrc/stress.cu``), and
``optimize_geometry`` drives the reference's optimisers with them.  Second derivatives (bulk modulus,
elastic constants, force constants: implicit differentiation through the density optimisation) are outside
this path and raise ``NotImplementedError``.
"""
import numpy as np
import torch

from .ion_utils import (get_ion_charge, interpolate_recpot, lattice_sum, ion_interaction_sum, ionic_potential,
                        ion_electron_forces, ion_electron_stress)
from .functional_tools import wavevecs
from ._optimizers.lbfgs.lbfgsnew import LBFGSNew
from ._optimizers.tpgd.two_point_gradient_descent import TPGD
from . import _density_opt


def _default_device():
    if not torch.cuda.is_available():
        raise RuntimeError('professad_b200.System needs a CUDA device (B200); none is visible and there is no '
                           'CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def _term_name(functional):
    return getattr(functional, '__qualname__', None) or getattr(functional, '__name__', '')


class System():
    """A periodic system for orbital-free DFT (system.py:18-72)."""

    # 2018 CODATA (system.py:26-33)
    m_per_bohr = 5.29177210903e-11
    A_per_b = m_per_bohr * 1e10
    J_per_Ha = 4.3597447222071e-18
    eV_per_Ha = J_per_Ha / 1.602176634e-19
    GPa_per_atomic = J_per_Ha / m_per_bohr**3 * 1e-9

    # set to False to force the generic host-driven optimiser even when every term is native
    use_native_optimizer = True

    def __init__(self, box_vecs, shape, ions, terms, units='b', coord_type='cartesian', Rc=None,
                 pme_order=None, device=None):
        self.__device = _default_device() if device is None else torch.device(device)
        self.__terms = terms
        self.__shape = tuple(int(s) for s in shape)
        self.__pme_order = pme_order
        self.__Rc = Rc
        self.__Eion_cache = None
        self.set_lattice(box_vecs, units, initialization=True)
        self.__process_ions(ions, coord_type, units)
        self.__update_ionic_potential()
        self.initialize_density()
        self.__ene = self.__compute_energy()

    @classmethod
    def ecut2shape(self, energy_cutoff, box_vecs):
        """Grid shape for an energy cutoff in eV and a lattice in Angstrom (system.py:74-89): always odd."""
        bvs = box_vecs / self.A_per_b
        kcut = np.sqrt(2 * energy_cutoff / self.eV_per_Ha)
        shape = 1 + 2 * torch.ceil(kcut / (2 * np.pi / torch.sqrt(torch.sum(bvs.pow(2), axis=1))))
        return tuple(shape.int().tolist())

    # ------------------------------------------------------------------ initialisation / updates
    def set_device(self, device=None):
        self.__device = _default_device() if device is None else torch.device(device)
        self.__box_vecs = self.__box_vecs.to(self.__device)
        self.__den = self.__den.to(self.__device)
        self.__v_ext = self.__v_ext.to(self.__device)
        self.__frac_ion_coords = self.__frac_ion_coords.to(self.__device)

    def __unit_factor(self, units):
        if units == 'a':
            return self.A_per_b
        if units == 'b':
            return 1.0
        raise ValueError('Parameter \'units\' can only be \'b\' (Bohr) or \'a\' (Angstrom)')

    def __process_ions(self, ions, coord_type, units):
        n_elec, ion_list, name, coords = 0, [], '', []
        for species in ions:   # [name, path_to_recpot, coordinates]
            charge = get_ion_charge(species[1])
            count = species[2].shape[0]
            ion_list.append((species[0], species[1], count, charge))
            coords.append(species[2].double().to(self.__device))
            n_elec += count * charge
            name += species[0] + str(int(count))
        ion_coords = torch.cat(coords) if coords else torch.empty((0, 3), dtype=torch.double, device=self.__device)
        self.__name = name
        self.__N_ions = ion_coords.shape[0]
        self.__N_elec = n_elec
        self.__ions = ion_list
        self.place_ions(ion_coords, coord_type, units, initialization=True)

    def place_ions(self, ion_coords, coord_type='cartesian', units='a', initialization=False):
        """system.py:125-157"""
        ion_coords = ion_coords.clone().double().to(self.__device)
        if coord_type == 'cartesian':
            frac = torch.matmul(ion_coords / self.__unit_factor(units), torch.linalg.inv(self.__box_vecs))
        elif coord_type == 'fractional':
            frac = ion_coords
        else:
            raise ValueError('Parameter \'coord_type\' can only be \'cartesian\' or \'fractional\'')
        frac = frac - torch.floor(frac)          # twice on purpose: -1e-17 -> 1.0 -> 0.0
        self.__frac_ion_coords = frac - torch.floor(frac)
        if not initialization:
            self.__update_ionic_potential()
            self.__ene = self.__compute_energy()

    def set_lattice(self, box_vecs, units='a', initialization=False):
        """system.py:159-181 : the density is rescaled to conserve the electron number."""
        factor = self.__unit_factor(units)
        if not initialization:
            old_vol = self.__vol()
        self.__box_vecs = box_vecs.clone().double().to(self.__device) / factor
        if not initialization:
            self.__update_ionic_potential()
            self.__den = self.__den * (old_vol / self.__vol())
            self.__ene = self.__compute_energy()

    def __species(self):
        out, first = [], 0
        for _, path, count, _ in self.__ions:
            out.append((path, self.__frac_ion_coords[first:first + count]))
            first += count
        return out

    def __potential_from_ions(self, cart_ion_coords):
        """system.py:183-205.  Exact structure factor (pme_order None, the reference's default): one native call
        that never materialises the N_k x N_ion phases; particle-mesh Ewald orders up to 32: native B-spline spreading
        + r2c (pad_ionic_potential_pme); higher orders: the torch spline path."""
        if self.__N_ions > 0 and (self.__pme_order is None or self.__pme_order <= 32):
            return ionic_potential(self.__box_vecs, self.__shape, self.__species(), self.__pme_order)
        kx, ky, kz, k2 = wavevecs(self.__box_vecs, self.__shape)
        k = torch.sqrt(k2)
        v_ext = torch.zeros(self.__shape, dtype=torch.double, device=self.__device)
        first = 0
        for _, path, count, _ in self.__ions:
            v_s_ft = interpolate_recpot(path, k)
            v_ext += lattice_sum(self.__box_vecs, self.__shape, cart_ion_coords[first:first + count], v_s_ft,
                                 self.__pme_order)
            first += count
        return v_ext

    def __update_ionic_potential(self):
        if any(_term_name(f) == 'IonElectron' for f in self.__terms):
            self.__v_ext = self.__potential_from_ions(torch.matmul(self.__frac_ion_coords, self.__box_vecs))
        else:
            self.__v_ext = torch.zeros(self.__shape, dtype=torch.double, device=self.__device)

    def set_potential(self, pot):
        assert tuple(pot.shape) == self.__shape, 'Shape of new potential must match the system\'s.'
        self.__v_ext = pot.clone().double().to(self.__device)
        self.__ene = self.__compute_energy()

    def initialize_density(self):
        """Uniform density N_elec / vol (system.py:218-222)."""
        self.__den = torch.full(self.__shape, float(self.__N_elec) / self.__vol().item(), dtype=torch.double,
                                device=self.__device)

    def set_density(self, den):
        assert tuple(den.shape) == self.__shape, 'Shape of new density must match the system\'s.'
        self.__den = den.double().to(self.__device)
        self.__ene = self.__compute_energy()

    def set_electron_number(self, N):
        self.__N_elec = N

    def __vol(self):
        return torch.abs(torch.linalg.det(self.__box_vecs))

    def detach(self):
        self.__box_vecs = self.__box_vecs.detach()
        self.__den = self.__den.detach()
        self.__v_ext = self.__v_ext.detach()
        self.__frac_ion_coords = self.__frac_ion_coords.detach()

    # ------------------------------------------------------------------------------- getters
    def device(self):
        return self.__device

    def name(self):
        return self.__name

    def ion_count(self):
        return self.__N_ions

    def electron_count(self):
        return self.__N_elec

    def lattice_vectors(self, units='a'):
        return self.__unit_factor(units) * self.__box_vecs

    def ions(self):
        return self.__ions

    def cartesian_ionic_coordinates(self, units='a'):
        return self.__unit_factor(units) * torch.matmul(self.__frac_ion_coords, self.__box_vecs)

    def fractional_ionic_coordinates(self):
        return self.__frac_ion_coords

    def ionic_potential(self, units='Ha'):
        if units == 'Ha':
            return self.__v_ext
        if units == 'eV':
            return self.__v_ext * self.eV_per_Ha
        raise ValueError('Parameter \'units\' can only be \'Ha\' or \'eV\'')

    def density(self, requires_grad=False):
        if requires_grad:
            self.__second_order('density(requires_grad=True)')
        return self.__den.detach()

    def volume(self, units='b3'):
        if units == 'b3':
            return self.__vol().item()
        if units == 'a3':
            return self.__vol().item() * self.A_per_b**3
        raise ValueError('Parameter \'units\' can only be \'b3\' or \'a3\'')

    def energy(self, units='Ha', requires_grad=False):
        if requires_grad:
            self.__second_order('energy(requires_grad=True)')
        E = self.__ene.item()
        if units == 'Ha':
            return E
        if units == 'eV':
            return E * self.eV_per_Ha
        raise ValueError('Parameter \'units\' can only be \'Ha\' or \'eV\'')

    # ------------------------------------------------------------- convergence measures / potentials
    def check_density_convergence(self, method='dEdchi'):
        """max |dE/dchi| or max |mu - dE/dn| (system.py:377-412)."""
        if method == 'dEdchi':
            return torch.max(torch.abs(self.functional_derivative('chi'))).item()
        elif method == 'euler':
            dEdn = self.functional_derivative('density')
            mu = torch.mean(dEdn * self.__den) * self.__vol() / self.__N_elec
            return torch.max(torch.abs(mu - dEdn)).item()

    def functional_derivative(self, type='density', requires_grad=False):
        """dE/dn or dE/dchi with n = N chi^2 / int chi^2 (system.py:414-447)."""
        if requires_grad:
            self.__second_order('functional_derivative(requires_grad=True)')
        self.detach()
        dV = self.__vol() / self.__den.numel()
        if type == 'density':
            den = self.__den.requires_grad_(True)
            E = self.__compute_energy(for_den_opt=True)
            dEdn = torch.autograd.grad(E, den)[0] / dV
            self.__den = self.__den.detach().requires_grad_(False)
            return dEdn
        elif type == 'chi':
            chi = torch.sqrt(self.__den).requires_grad_(True)
            N_tilde = torch.mean(chi.pow(2)) * self.__vol()
            self.__den = (self.__N_elec / N_tilde) * chi.pow(2)
            E = self.__compute_energy(for_den_opt=True)
            dEdchi = torch.autograd.grad(E, chi)[0] / dV
            self.__den = self.__den.detach()
            return dEdchi

    def chemical_potential(self):
        dEdn = self.functional_derivative('density')
        return (torch.mean(dEdn * self.__den) * self.__vol() / self.__N_elec).item()

    # ------------------------------------------------------------------------- out-of-scope API
    def __second_order(self, what):
        raise NotImplementedError(f'System.{what}: implicit differentiation / second derivatives are outside the '
                                  'B200 hot path (SURVEY.md section 8: out of scope)')

    def __pressure_units(self, units):
        if units == 'Ha/b3':
            return 1.0
        if units == 'eV/a3':
            return self.eV_per_Ha / self.A_per_b**3
        if units == 'GPa':
            return self.GPa_per_atomic
        raise ValueError('Parameter \'units\' can only be \'Ha/b3\', \'eV/a3\' or \'GPa\'')

    def pressure(self, units='Ha/b3', requires_grad=False):
        """P = -dE/dvol at fixed electron number and fractional ionic coordinates (system.py:494-522)
        = -trace(stress) / 3."""
        if requires_grad:
            self.__second_order('pressure(requires_grad=True)')
        factor = self.__pressure_units(units)
        return -torch.trace(self.__compute_stress()).item() / 3 * factor

    def enthalpy(self, units='Ha'):
        """H = E + P vol (system.py:524-540)."""
        H = self.energy('Ha') + self.pressure('Ha/b3') * self.volume('b3')
        if units == 'Ha':
            return H
        if units == 'eV':
            return H * self.eV_per_Ha
        raise ValueError('Parameter \'units\' can only be \'Ha\' or \'eV\'')

    def bulk_modulus(self, units='Ha/b3', requires_grad=False):
        self.__second_order('bulk_modulus')

    def __native_pme_order(self):
        """pme_order for the native force / stress kernels (orders the B-spline kernels cover), else None = exact."""
        return self.__pme_order if (self.__pme_order is not None and self.__pme_order <= 32) else None

    def forces(self, units='Ha/b'):
        """F = -dE/dR at fixed density (system.py:623-643, 913-925): the IonElectron part is one native
        reciprocal-space reduction per ion (pad_ion_forces) or, when the System was built with ``pme_order``, the
        derivative of the particle-mesh structure factor as in the reference (pad_ion_forces_pme: one c2r and a gather of
        B-spline derivative weights); the IonIon part has closed forms (pad_ion_ion)."""
        if units not in ('Ha/b', 'eV/a'):
            raise ValueError('Parameter \'units\' can only be \'Ha/b\' or \'eV/a\'')
        names = [_term_name(f) for f in self.__terms]
        forces = torch.zeros((self.__N_ions, 3), dtype=torch.double, device=self.__device)
        if 'IonElectron' in names:
            forces = forces + ion_electron_forces(self.__box_vecs, self.__den, self.__species(), self.__native_pme_order())
        if 'IonIon' in names:
            cart = torch.matmul(self.__frac_ion_coords, self.__box_vecs).detach().requires_grad_(True)
            U = self.__ion_ion_interaction(cart)
            forces = forces - torch.autograd.grad(U, cart)[0]
        return forces if units == 'Ha/b' else forces * self.eV_per_Ha / self.A_per_b

    def __compute_stress(self):
        """system.py:927-935: (1/vol) dE/d eps with the density rescaled to conserve N and the ions at fixed
        fractional coordinates, symmetrised.  Native terms: analytic kernels (pad_stress_terms, pad_ion_stress);
        IonIon: autograd through the real-space pair sum."""
        plain = [f for f in self.__terms if _term_name(f) not in ('IonElectron', 'IonIon')]
        names = [_term_name(f) for f in self.__terms]
        sig = torch.zeros((3, 3), dtype=torch.double, device=self.__device)
        if plain:
            T = _density_opt.describe_terms(plain, self.__device)
            if T is None:
                raise NotImplementedError('System.stress: analytic stresses exist for the native functionals only '
                                          '(user-defined Python terms need autograd through box_vecs)')
            sig = sig + _density_opt.stress_terms(self.__box_vecs, self.__den, T)
        if 'IonElectron' in names:
            sig = sig + ion_electron_stress(self.__box_vecs, self.__den, self.__species(), self.__native_pme_order())
        if 'IonIon' in names:
            box = self.__box_vecs.detach().clone().requires_grad_(True)
            saved, self.__box_vecs = self.__box_vecs, box
            try:
                U = self.__ion_ion_interaction(torch.matmul(self.__frac_ion_coords, box))
            finally:
                self.__box_vecs = saved
            dEdcell = torch.autograd.grad(U, box)[0].T
            sig = sig + torch.matmul(dEdcell, self.__box_vecs) / self.__vol()
        return 0.5 * (sig + sig.T)

    def stress(self, units='Ha/b3'):
        """Stress tensor (system.py:645-668)."""
        return self.__compute_stress() * self.__pressure_units(units)

    def elastic_constants(self, units='Ha/b3'):
        self.__second_order('elastic_constants')

    def force_constants(self, primitive_ion_indices, units='eV/a2'):
        self.__second_order('force_constants')

    # -------------------------------------------------------------------------- geometry optimisation
    def optimize_geometry(self, ftol=0.02, stol=0.002, g_conv_cond_count=3, g_method='LBFGSlinesearch',
                          g_step_size=0.1, g_maxiter=1000, g_verbose=False, **den_opt_kwargs):
        """Minimise the energy over the fractional ionic coordinates (``ftol`` in eV/A, None = ions fixed) and / or
        the lattice vectors (``stol`` in eV/A^3, None = cell fixed); system.py:937-1068.  Same optimisers, closure
        and stop rule; the gradients come from the analytic forces and stress instead of autograd.  ``g_max_step``
        (extra keyword, default 1.0; None = the reference's unguarded behaviour): L-BFGS restart guard, see LBFGSNew --
        a move of more than one whole cell (fractional coordinates) or 1 bohr (lattice-vector components) per inner
        iteration is replaced by a steepest-descent restart."""
        den_opt_kwargs.setdefault('g_max_step', 1.0)
        if (ftol is None) and (stol is None):
            raise ValueError('At least one of \'stol\' or \'ftol\' cannot be \'None\'')
        n_f = 3 * self.__N_ions if ftol is not None else 0
        frac0, box0 = self.__frac_ion_coords.detach().clone(), self.__box_vecs.detach().clone()
        pieces = ([frac0.reshape(-1)] if ftol is not None else []) + ([box0.reshape(-1)] if stol is not None else [])
        params = torch.cat(pieces).clone()

        def geometry(p):
            frac = p[:n_f].reshape(-1, 3) if ftol is not None else frac0
            box = p[n_f:].reshape(3, 3) if stol is not None else box0
            return box, frac
        return self.optimize_parameterized_geometry(params, geometry, ftol, stol, g_conv_cond_count, g_method,
                                                    g_step_size, g_maxiter, g_verbose, None, **den_opt_kwargs)

    def optimize_parameterized_geometry(self, params, parameterized_geometry, ftol=0.02, stol=0.002,
                                        g_conv_cond_count=3, g_method='LBFGSlinesearch', g_step_size=0.1,
                                        g_maxiter=1000, g_verbose=False, param_string=None, **den_opt_kwargs):
        """system.py:1070-1198: ``parameterized_geometry(params) -> (box_vecs [bohr], frac_ion_coords)``.
        Closure (system.py:1124-1137): at fixed chi = sqrt(n) the ions and the cell move, v_ext is rebuilt, the
        density is renormalised to N electrons, E is the total energy.  dE/dparams is assembled from the analytic
        pieces -- dE/dfrac = -F box^T, dE/dbox = vol box^-T sigma -- and pulled back through the user's
        parametrisation by autograd."""
        den_opt_inputs = {'ntol': 1e-10, 'n_conv_cond_count': 3, 'n_method': 'LBFGS', 'n_step_size': 0.1,
                          'n_maxiter': 1000, 'conv_target': 'dE', 'n_verbose': False, 'from_uniform': False}
        g_max_step = den_opt_kwargs.pop('g_max_step', None)
        den_opt_inputs.update(den_opt_kwargs)
        if (ftol is None) and (stol is None):
            raise ValueError('At least one of \'stol\' or \'ftol\' cannot be \'None\'')
        params = params.detach().to(device=self.__device, dtype=torch.double).clone().requires_grad_(True)
        if g_method == 'RPROP':
            optimizer = torch.optim.Rprop([params], lr=g_step_size)
        elif g_method == 'TPGD':
            optimizer = TPGD([params], lr=g_step_size)
        elif g_method == 'LBFGSlinesearch':
            optimizer = LBFGSNew([params], lr=g_step_size, history_size=8, max_iter=6, line_search_fn=True, max_step=g_max_step)
        elif g_method == 'LBFGS':
            optimizer = LBFGSNew([params], lr=g_step_size, history_size=8, max_iter=6, max_step=g_max_step)
        else:
            raise ValueError('Only \'LBFGSlinesearch\', \'LBFGS\', \'RPROP\' or \'TPGD\' recognized for \'g_method\'')
        state = {'chi': None}

        def closure():
            if torch.is_grad_enabled():
                optimizer.zero_grad()
            with torch.enable_grad():
                box, frac = parameterized_geometry(params)
            self.__box_vecs = box.detach().to(self.__device).double()
            self.__frac_ion_coords = frac.detach().to(self.__device).double()
            self.__update_ionic_potential()
            chi = state['chi']
            N_tilde = torch.mean(chi.pow(2)) * self.__vol()
            self.__den = (self.__N_elec / N_tilde) * chi.pow(2)
            E = self.__compute_energy()
            outs, grads = [], []
            if frac.requires_grad:
                outs.append(frac)
                grads.append(-torch.matmul(self.forces('Ha/b'), self.__box_vecs.T).to(frac.dtype))
            if box.requires_grad:
                outs.append(box)
                grads.append(self.__vol() * torch.matmul(torch.linalg.inv(self.__box_vecs).T, self.__compute_stress()))
            params.grad = torch.autograd.grad(outs, params, grads)[0] if outs else torch.zeros_like(params)
            return E.detach()

        def max_abs(fn, *a):
            try:
                return torch.max(torch.abs(fn(*a))).item()
            except NotImplementedError:
                return float('nan')

        self.optimize_density(**den_opt_inputs)
        E_prev = self.energy('eV') / self.ion_count()
        head = '{:^7} {:^20} {:^20} {:^20} {:^20}'.format('Iter', 'E [eV per atom]', 'dE [eV per atom]', 'Max Force [eV/Å]',
                                                          'Max Stress [eV/Å³]')
        row = '{:^7} {:^20.6f} {:^20.6g} {:^20.6g} {:^20.6g}'
        if g_verbose:
            print(head + ('Params' if param_string is not None else ''), flush=True)
            print(row.format(0, E_prev, 0, max_abs(self.forces, 'eV/a'), max_abs(self.stress, 'eV/a3'))
                  + (param_string(params) if param_string is not None else ''), flush=True)
        conv_counter, success_iter = 0, None
        self.last_geometry_optimization = {'iterations': 0, 'converged': False}
        for it in range(1, round(g_maxiter) + 1):
            state['chi'] = torch.sqrt(self.__den)
            optimizer.step(closure)
            # The optimiser's last move is not followed by a closure: bring the System to the final parameters.  (The
            # reference's System sees the moved tensors too, but re-optimises the density in the ionic potential of the
            # last closure, system.py:1024-1027; here v_ext is rebuilt so that E, forces and stress of an iteration
            # all belong to the same geometry.)
            with torch.no_grad():
                box, frac = parameterized_geometry(params.detach())
            self.__box_vecs = box.detach().to(self.__device).double().clone()
            self.__frac_ion_coords = frac.detach().to(self.__device).double().clone()
            self.__Eion_cache = None        # the cached ion-ion energy belongs to the last closure's (line-search) geometry
            self.__update_ionic_potential()
            self.detach()
            self.optimize_density(**den_opt_inputs)
            E_new = self.energy('eV') / self.ion_count()
            max_force = max_abs(self.forces, 'eV/a')
            max_stress = max_abs(self.stress, 'eV/a3') if (stol is not None or g_verbose) else float('nan')
            if g_verbose:
                print(row.format(it, E_new, E_new - E_prev, max_force, max_stress)
                      + (param_string(params) if param_string is not None else ''), flush=True)
            E_prev = E_new
            if it > 3:      # convergence is only checked after the 3rd iteration
                ok = (ftol is None or max_force < ftol) and (stol is None or max_stress < stol)
                conv_counter = conv_counter + 1 if ok else 0
            self.last_geometry_optimization = {'iterations': it, 'converged': False, 'max_force_eV_A': max_force,
                                               'max_stress_eV_A3': max_stress, 'energy_eV_per_atom': E_new}
            if conv_counter == g_conv_cond_count:
                success_iter = it
                self.last_geometry_optimization['converged'] = True
                break
        if g_verbose:
            if success_iter is not None:
                print('Geometry optimization successfully converged in {} step(s) \n'.format(success_iter), flush=True)
            else:
                print('Geometry optimization failed to converge in {} step(s) \n'.format(g_maxiter), flush=True)
        return success_iter is not None

    # ------------------------------------------------------------------------------- ion-ion
    def set_Rc(self, Rc=None):
        self.__Rc = Rc

    def __ion_ion_interaction(self, cart_ion_coords):
        """system.py:733-754 : R_d = 2 h_max, R_c = 3 R_d^2 / h_max unless Rc is given."""
        charges = torch.cat([torch.full((count,), float(z), dtype=torch.double, device=self.__device)
                             for _, _, count, z in self.__ions])
        h_max = torch.max(1 / torch.sqrt(torch.sum(torch.linalg.inv(self.__box_vecs.detach().T).pow(2), 1)))
        if self.__Rc is None:
            Rd = 2 * h_max
            Rc = 3 * Rd * Rd / h_max
        else:
            Rc = self.__Rc
            Rd = torch.sqrt(h_max * Rc / 3)
        E_ion = ion_interaction_sum(self.__box_vecs, cart_ion_coords, charges, Rc, Rd)
        self.__Eion_cache = E_ion.item()
        return E_ion

    # --------------------------------------------------------------------- energy and optimisation
    def __compute_energy(self, for_den_opt=False, use_ion_cache=False):
        """Sum of the terms (system.py:759-772); IonElectron gets v_ext, IonIon is skipped inside the
        density optimisation."""
        E = torch.zeros((1,), dtype=torch.double, device=self.__device)
        for functional in self.__terms:
            name = _term_name(functional)
            if name == 'IonElectron':
                E = E + functional(self.__box_vecs, self.__den, self.__v_ext)
            elif name == 'IonIon':
                if not for_den_opt:
                    if use_ion_cache and self.__Eion_cache is not None:
                        E = E + self.__Eion_cache
                    else:
                        E = E + self.__ion_ion_interaction(torch.matmul(self.__frac_ion_coords, self.__box_vecs))
            else:
                E = E + functional(self.__box_vecs, self.__den)
        return E

    def optimize_density(self, ntol=1e-7, n_conv_cond_count=3, n_method='LBFGS', n_step_size=0.1,
                         n_maxiter=1000, conv_target='dE', n_verbose=False, from_uniform=False,
                         potentials=None):
        """Direct minimisation over chi = sqrt(n) (system.py:774-908); same arguments and stop rule:
        after the 5th iteration, ``conv_target`` below ``ntol`` on ``n_conv_cond_count`` consecutive
        iterations."""
        if n_method not in ('LBFGS', 'TPGD'):
            raise ValueError('Only \'LBFGS\' or \'TPGD\' recognized for \'n_method\' argument')
        if conv_target not in ('dE', 'dEdchi', 'euler'):
            raise ValueError('Only \'dE\', \'dEdchi\' or \'euler\' recognized as \'conv_target\' argument')
        self.detach()
        if from_uniform:
            self.initialize_density()
        else:
            current_den = self.__den
            current_E = self.__compute_energy(for_den_opt=True)
            self.initialize_density()
            uniform_E = self.__compute_energy(for_den_opt=True)
            if current_E < uniform_E:
                self.set_density(current_den)

        T = (_density_opt.describe_terms(self.__terms, self.__device)
             if (potentials is None and self.use_native_optimizer) else None)
        if T is not None:
            self.__optimize_density_native(T, ntol, n_conv_cond_count, n_method, n_step_size, n_maxiter, conv_target,
                                           n_verbose)
            return

        chi = torch.sqrt(self.__den).requires_grad_()
        if n_method == 'LBFGS':
            optimizer = LBFGSNew([chi], lr=n_step_size, history_size=8, max_iter=6)
        else:
            optimizer = TPGD([chi], lr=n_step_size)
        vol, dV = self.__vol(), self.__vol() / self.__den.numel()

        if potentials is None:
            def closure():
                if torch.is_grad_enabled():
                    optimizer.zero_grad()
                N_tilde = torch.mean(chi.pow(2)) * vol
                self.__den = (self.__N_elec / N_tilde) * chi.pow(2)
                E = self.__compute_energy(for_den_opt=True)
                if E.requires_grad:
                    E.backward()
                return E
        else:
            def closure():
                if torch.is_grad_enabled():
                    optimizer.zero_grad()
                chi.requires_grad = False       # user potentials may use autograd themselves
                N_tilde = torch.mean(chi.pow(2)) * vol
                self.__den = (self.__N_elec / N_tilde) * chi.pow(2)
                E = self.__compute_energy(for_den_opt=True)
                dEdn = potentials(self.__box_vecs, self.__den)
                mu = torch.mean(dEdn * self.__den) * vol / self.__N_elec
                grad = (self.__N_elec / N_tilde) * 2 * chi * (dEdn - mu) * dV
                chi.requires_grad = True
                chi.grad = grad
                return E

        E_prev = self.__compute_energy(for_den_opt=True).item() * self.eV_per_Ha
        if n_verbose:
            print('Starting density optimization')
            print('{:^8} {:^12} {:^12} {:^18} {:^18}'.format('Iter', 'E [eV]', 'dE [eV]', 'Max |𝛿E/𝛿χ|', 'Max |µ-𝛿E/𝛿n|'))
            print('{:^8} {:^12.6f} {:^12.6g} {:^18.6g} {:^18.6g}'.format(
                0, E_prev, 0, self.check_density_convergence('dEdchi'), self.check_density_convergence('euler')))

        conv_counter = 0
        self.last_optimization = {'iterations': 0, 'converged': False}
        for it in range(1, round(n_maxiter) + 1):
            optimizer.step(closure)
            dEdchi = torch.abs(chi.grad / dV).max().item()
            with torch.no_grad():
                E = self.__compute_energy(for_den_opt=True).item() * self.eV_per_Ha
            dE, E_prev = E - E_prev, E
            if n_verbose or conv_target == 'euler':
                euler = self.check_density_convergence('euler')
            if n_verbose:
                print('{:^8} {:^12.6f} {:^12.6g} {:^18.6g} {:^18.6g}'.format(it, E_prev, dE, dEdchi, euler))
            stop_var = {'dE': abs(dE), 'dEdchi': dEdchi}.get(conv_target)
            if conv_target == 'euler':
                stop_var = euler
            if it > 5:
                conv_counter = conv_counter + 1 if stop_var < ntol else 0
            self.last_optimization['iterations'] = it
            if conv_counter == n_conv_cond_count:
                self.last_optimization['converged'] = True
                if n_verbose:
                    print('Density optimization successfully converged in {} step(s) \n'.format(it))
                break
            if it == round(n_maxiter) and n_verbose:
                print('Density optimization failed to converge in {} steps \n'.format(int(it)))
        self.detach()
        self.__ene = self.__compute_energy(use_ion_cache=True)

    def __optimize_density_native(self, T, ntol, n_conv_cond_count, n_method, n_step_size, n_maxiter, conv_target,
                                  n_verbose):
        """Every term is native: run the whole loop on the device (pad_denopt_run).  Same iterates and
        stop rule as the generic path; the per-iteration table is printed once the loop has finished
        because the host never waits on an individual iteration."""
        if n_verbose:
            print('Starting density optimization')
            print('{:^8} {:^12} {:^12} {:^18} {:^18}'.format('Iter', 'E [eV]', 'dE [eV]', 'Max |𝛿E/𝛿χ|', 'Max |µ-𝛿E/𝛿n|'))
            E0 = self.__compute_energy(for_den_opt=True).item() * self.eV_per_Ha
            print('{:^8} {:^12.6f} {:^12.6g} {:^18.6g} {:^18.6g}'.format(
                0, E0, 0, self.check_density_convergence('dEdchi'), self.check_density_convergence('euler')))
        den = self.__den.detach().clone().contiguous()
        res, trace = _density_opt.run(self.__box_vecs, den, self.__v_ext, T, self.__N_elec, ntol, n_conv_cond_count,
                                      n_method, n_step_size, n_maxiter, conv_target)
        self.__den = den
        self.last_optimization = dict(res, trace=trace, native=True)
        if n_verbose:
            for i, row in enumerate(trace.tolist(), start=1):
                print('{:^8} {:^12.6f} {:^12.6g} {:^18.6g} {:^18.6g}'.format(i, *row))
            if res['converged']:
                print('Density optimization successfully converged in {} step(s) \n'.format(res['iterations']))
            else:
                print('Density optimization failed to converge in {} steps \n'.format(res['iterations']))
        self.detach()
        self.__ene = self.__compute_energy(use_ion_cache=True)

    # ------------------------------------------------------------------------------ strain scan
    def eos_fit(self, f=0.05, N=9, eos='bm', verbose=False, plot=False, **den_opt_kwargs):
        """Energy-volume scan + equation-of-state fit (system.py:568-621).  Returns (params, err) in
        the reference's order: K0 [GPa], K0', E0 [eV], V0 [A^3]."""
        from .elastic_tools import fit_eos
        opts = {'ntol': 1e-10, 'n_conv_cond_count': 3, 'n_method': 'LBFGS', 'n_step_size': 0.1, 'n_maxiter': 1000,
                'conv_target': 'dE', 'n_verbose': False, 'from_uniform': False}
        opts.update(den_opt_kwargs)
        v0 = self.volume('a3')
        shape_vecs = self.lattice_vectors('a') / v0**(1 / 3)
        volumes, energies = [], []
        if verbose:
            print('\n{:^22} {:^22}'.format('Volume [Å³ per atom]', 'Energy [eV per atom]'))
        for v in v0 * np.linspace(1 - f, 1 + f, N):
            self.set_lattice(v**(1 / 3) * shape_vecs, units='a')
            self.optimize_density(**opts)
            volumes.append(self.volume('a3') / self.__N_ions)
            energies.append(self.energy('eV') / self.__N_ions)
            if verbose:
                print('{:^22.10f} {:^22.10f}'.format(volumes[-1], energies[-1]))
        params, err = fit_eos(volumes, energies, eos, plot)
        to_gpa = self.GPa_per_atomic / (self.eV_per_Ha / self.A_per_b**3)
        params[0] *= to_gpa
        err[0] *= to_gpa
        return params, err
