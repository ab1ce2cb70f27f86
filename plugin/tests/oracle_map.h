/* run-check on a machine without a GPU: the adapter's calls into the C-ABI go to the CPU oracle, which exports the same
 * entry points under the prefix snb_oracle_ (test infrastructure; force-included by plugin/Makefile, never by the plugin). */
#ifndef SNB_ORACLE_MAP_H_
#define SNB_ORACLE_MAP_H_
#define snb_create snb_oracle_create
#define snb_destroy snb_oracle_destroy
#define snb_execute snb_oracle_execute
#define snb_update_parameters snb_oracle_update_parameters
#define snb_get_pme_parameters snb_oracle_get_pme_parameters
#define snb_get_ljpme_parameters snb_oracle_get_ljpme_parameters
#define snb_last_error snb_oracle_last_error
#endif
