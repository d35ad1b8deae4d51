"""SCML plugin API of the reference (physical_systems/__init__.py) for the device path."""
from .converters import (ContB6BridgeConverter, ContDynamicallyAveragedConverter, ContFourQuadrantConverter, ContMultiConverter,
                         ContOneQuadrantConverter, ContTwoQuadrantConverter, FiniteB6BridgeConverter, FiniteConverter,
                         FiniteFourQuadrantConverter, FiniteMultiConverter, FiniteOneQuadrantConverter, FiniteTwoQuadrantConverter,
                         PowerElectronicConverter)
from .electric_motors import (DcExternallyExcitedMotor, DcMotor, DcPermanentlyExcitedMotor, DcSeriesMotor, DcShuntMotor, DoublyFedInductionMotor, ElectricMotor,
                              ExternallyExcitedSynchronousMotor, InductionMotor, PermanentMagnetSynchronousMotor,
                              SquirrelCageInductionMotor, SynchronousMotor, SynchronousReluctanceMotor, ThreePhaseMotor)
from .mechanical_loads import ConstantSpeedLoad, ExternalSpeedLoad, MechanicalLoad, OrnsteinUhlenbeckLoad, PolynomialStaticLoad
from .physical_systems import (DcMotorSystem, DoublyFedInductionMotorSystem, ExternallyExcitedSynchronousMotorSystem, PhysicalSystem,
                               SCMLSystem, SquirrelCageInductionMotorSystem, SynchronousMotorSystem, ThreePhaseMotorSystem)
from .solvers import EulerSolver, OdeSolver, RK4Solver, ScipyOdeIntSolver, ScipyOdeSolver, ScipySolveIvpSolver
from .voltage_supplies import AC1PhaseSupply, AC3PhaseSupply, IdealVoltageSupply, RCVoltageSupply, VoltageSupply

# short names the reference exports for the converters (physical_systems/__init__.py:14-24)
Cont1QC, Cont2QC, Cont4QC, ContB6C, ContMulti = (ContOneQuadrantConverter, ContTwoQuadrantConverter, ContFourQuadrantConverter, ContB6BridgeConverter,
                                                  ContMultiConverter)
Finite1QC, Finite2QC, Finite4QC, FiniteB6C, FiniteMulti = (FiniteOneQuadrantConverter, FiniteTwoQuadrantConverter, FiniteFourQuadrantConverter,
                                                            FiniteB6BridgeConverter, FiniteMultiConverter)
