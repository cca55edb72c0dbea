/*
 * vlist_select_stub.h -- placed by oracle/Makefile in the build's "view" directory under the name
 * n2000_k1000_no6cycle_ldpc_Vlist_device.h, the table header ldpc_erasure_decoder.cl:22,
 * ldpc_erasure_decoder_old.pro:22 and ldpc_erasure_encoder.cl:18 include.
 *
 * Why a stand-in: as committed, that header (a) has n_ldpc / k_ldpc commented out (:12-13), so the
 * datapath files do not compile against it, (b) redefines ldpc_params next to LDPC_Vlist_data.h, which
 * ldpc_erasure_decoder_top.cl:34 includes too, and (c) holds only the (2000,1000) rows.  The stand-in
 * selects a code's rows out of the reference's own master table (LDPC_Vlist_data.h:20) exactly as the
 * reference's perf_tests variant does at run time (ldpc_erasure_decoder_perf_tests.cl:33-43): rows
 * ldpc_params[code][2] .. ldpc_params[code][3].  No table data lives here; tests check that rows 0-999 of
 * the master table equal the (2000,1000) header this file stands in for.
 */
#ifndef REF_VLIST_SELECT_STUB_H
#define REF_VLIST_SELECT_STUB_H
static __thread int ref_code_ind;
#define n_ldpc (ldpc_params[ref_code_ind][0])
#define k_ldpc (ldpc_params[ref_code_ind][1])
#define parity_check_mat_Vlist (parity_check_mat_Vlist_master + ldpc_params[ref_code_ind][2])
#endif
