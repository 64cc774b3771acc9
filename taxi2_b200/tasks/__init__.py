"""Task classes built on the batch engine: the callers either side of the hot path
(/root/reference/src/itaxotools/taxi2/tasks/).  Same attribute surface (`input`, `params`,
`work_dir`, `progress_handler`, `start() -> Results`) and the same output files; the per-pair
generator chains are replaced by tile-wise device launches whose results are then written in the
reference's order."""
from .common import Results, console_report  # noqa: F401
from .decontaminate import Decontaminate, Decontaminate2  # noqa: F401
from .dereplicate import Dereplicate  # noqa: F401
from .versus_all import VersusAll  # noqa: F401
from .versus_reference import VersusReference  # noqa: F401
