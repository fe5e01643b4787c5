from .general_utils import *  # noqa: F401,F403  (the reference does `from ...utils import *`)
from .general_utils import (METAPATHS, metapath_table, update_pea_graph_input, get_folder_path, get_opt_class,
                            save_model, load_model, remap_legacy_state_dict, save_global_logger,
                            load_global_logger, load_dataset, instantwrite, clearcache)
