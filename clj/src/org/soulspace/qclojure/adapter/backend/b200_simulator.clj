(ns org.soulspace.qclojure.adapter.backend.b200-simulator
  "B200 state-vector backends for QClojure: `QuantumBackend` / `MultiDeviceBackend` records
  (application/backend.clj:72-131) whose execution path is libqcb200.so (include/qcb200.h) instead of
  domain/circuit/execute-circuit and the per-shot Clojure trajectory loop.

    B200Simulator          mirrors LocalQuantumSimulator      (adapter/backend/ideal_simulator.clj:100-176)
    B200HardwareSimulator  mirrors QuantumHardwareSimulator   (adapter/backend/hardware_simulator.clj:281-391)

  NOT COMPILED OR RUN IN THIS REPOSITORY'S CI: the build image has no JVM (DESIGN.md section 1).  What CAN be checked
  without one is checked: tests/test_clj_shim_layout.py parses `struct-layouts` and `kind-code` below and compares every
  offset, size and enum value with offsetof / sizeof / the enum of include/qcb200.h as compiled by gcc, and checks that
  every C symbol named in an `ffi` form is exported by libqcb200.so.  The same C entry points, in the same order, are
  exercised on the GPU through the Python mirror (qclojure_b200/backend.py).

  Result extraction stays QClojure's own code: for states up to `:max-state-qubits` (default 20) the final state comes
  back as a QClojure state map and `domain.result/extract-results` runs on it unchanged; only the measurement shots are
  taken from the device sampler (same measure-state rule, state.clj:894-913).  Larger states stay on the GPU and
  answer the specs the device can evaluate in place (:measurements, :hamiltonian).

  Binding: Java FFM (JDK 22+, java.lang.foreign); JNA variant for older JVMs at the bottom.
  Every C function returns int32 status; 0 = ok, message via qcb_last_error."
  (:require [org.soulspace.qclojure.application.backend :as backend]
            [org.soulspace.qclojure.application.hardware-optimization :as hwopt]
            [org.soulspace.qclojure.domain.channel :as channel]
            [org.soulspace.qclojure.domain.circuit :as circuit]
            [org.soulspace.qclojure.domain.operation-registry :as opreg]
            [org.soulspace.qclojure.domain.result :as result]
            [org.soulspace.qclojure.domain.state :as state]
            [fastmath.complex :as fc])
  (:import (java.lang.foreign Arena FunctionDescriptor Linker MemoryLayout MemorySegment SymbolLookup ValueLayout)
           (java.lang.invoke MethodHandle)))

;;; ------------------------------------------------------------------ FFM plumbing
(def ^:private ^Linker linker (Linker/nativeLinker))
(def ^:private lib-arena (Arena/global))
(defonce ^:private lookup
  (delay (SymbolLookup/libraryLookup (or (System/getProperty "qcb200.lib") "libqcb200.so") lib-arena)))

(def ^:private I32 ValueLayout/JAVA_INT)
(def ^:private I64 ValueLayout/JAVA_LONG)
(def ^:private F64 ValueLayout/JAVA_DOUBLE)
(def ^:private PTR ValueLayout/ADDRESS)

(defn- ^MethodHandle ffi
  "Downcall handle for `int32_t name(args...)`."
  [name & arg-layouts]
  (.downcallHandle linker (.orElseThrow (.find ^SymbolLookup @lookup name))
                   (FunctionDescriptor/of I32 (into-array MemoryLayout arg-layouts))))

(defn- call [handle-delay & args]
  (int (.invokeWithArguments ^MethodHandle @handle-delay ^java.util.List (vec args))))

;; include/qcb200.h - one downcall handle per entry point this shim binds
(def ^:private qcb-abi-version   (delay (ffi "qcb_abi_version")))
(def ^:private qcb-device-count  (delay (ffi "qcb_device_count" PTR)))
(def ^:private qcb-config-default (delay (ffi "qcb_config_default" PTR)))
(def ^:private qcb-create        (delay (ffi "qcb_create" PTR PTR)))                 ; (const qcb_config*, qcb_handle*)
(def ^:private qcb-destroy       (delay (ffi "qcb_destroy" PTR)))
(def ^:private qcb-last-error    (delay (ffi "qcb_last_error" PTR PTR I64)))
(def ^:private qcb-submit        (delay (ffi "qcb_submit" PTR PTR PTR)))             ; (h, const qcb_job_request*, uint64* id)
(def ^:private qcb-job-status    (delay (ffi "qcb_job_status" PTR I64 PTR)))
(def ^:private qcb-job-result    (delay (ffi "qcb_job_result_get" PTR I64 PTR)))
(def ^:private qcb-job-release   (delay (ffi "qcb_job_release" PTR I64)))
(def ^:private qcb-cancel        (delay (ffi "qcb_cancel" PTR I64 PTR)))
(def ^:private qcb-set-zero      (delay (ffi "qcb_set_zero" PTR)))
(def ^:private qcb-get-state     (delay (ffi "qcb_get_state" PTR I64 I64 PTR)))
(def ^:private qcb-noisy-set-initial-state (delay (ffi "qcb_noisy_set_initial_state" PTR PTR I64)))
(def ^:private qcb-noisy-draws-per-shot (delay (ffi "qcb_noisy_draws_per_shot" PTR PTR I64 PTR PTR)))
(def ^:private qcb-run-noisy     (delay (ffi "qcb_run_noisy" PTR PTR I64 PTR PTR I64 I64 PTR PTR I64)))

;;; ------------------------------------------------------------------ struct layouts (checked against gcc by the test-suite)
(def struct-layouts
  "Byte offsets and sizes of the C structs this shim fills by hand (include/qcb200.h, ABI version 2)."
  {:qcb_config       {:size 104 :n_qubits 0 :device 4 :fusion 8 :strict_parity 12 :tile_bits 16 :low_bits 20 :rank 24
                      :world_size 28 :nccl_unique_id 32 :max_stage_cost 40 :max_stage_rounds 44 :dense_mma 48
                      :tile_mover 52 :n_gpus 56 :device_ids 60 :reserved 92}
   :qcb_op           {:size 112 :kind 0 :q 4 :n_mask 16 :mask 24 :angle 32 :mat 40 :ext 104}
   :qcb_job_request  {:size 80 :ops 0 :n_ops 8 :initial_state 16 :initial_count 24 :uniforms 32 :n_shots 40
                      :ham_coeffs 48 :ham_strings 56 :n_terms 64 :want_probabilities 72 :want_state 76}
   :qcb_job_result   {:size 336 :status 0 :execution_time_ms 8 :n_shots 16 :outcomes 24 :energy 32 :has_energy 40
                      :probabilities 48 :prob_capacity 56 :state 64 :state_capacity 72 :error_message 80}
   :qcb_noise_entry  {:size 264 :op_kind 0 :n_kraus 4 :kraus 8}
   :qcb_noise_table  {:size 40 :entries 0 :n_entries 8 :has_readout 12 :prob_0_to_1 16 :prob_1_to_0 24 :correlation 32}})

(defn- off ^long [s field] (long (get-in struct-layouts [s field])))
(defn- size-of ^long [s] (long (get-in struct-layouts [s :size])))

(def kind-code
  "enum qcb_op_kind (include/qcb200.h) for every branch of apply-gate-to-state (domain/circuit.clj:964-1071) plus :measure."
  {:i 0 :x 1 :y 2 :z 3 :h 4 :s 5 :s-dag 6 :t 7 :t-dag 8 :rx 9 :ry 10 :rz 11 :phase 12 :cnot 13 :cz 14 :cy 15
   :crx 16 :cry 17 :crz 18 :swap 19 :iswap 20 :toffoli 21 :fredkin 22 :rydberg-cz 23 :rydberg-cphase 24
   :rydberg-blockade 25 :global-h 26 :global-x 27 :global-y 28 :global-z 29 :global-rx 30 :global-ry 31 :global-rz 32
   :measure 39})

;;; ------------------------------------------------------------------ circuit map -> qcb_op[]
(defn- operands
  "The operand order the C ABI expects, from the reference's :operation-params keys (circuit.clj:170-845).
  A missing :target defaults to qubit 0 exactly like circuit.clj:965-984."
  [op-type p]
  (case op-type
    (:cnot :cz :cy :crx :cry :crz :rydberg-cz :rydberg-cphase) [(:control p) (:target p) -1]
    (:swap :iswap) [(:qubit1 p) (:qubit2 p) -1]
    :toffoli [(:control1 p) (:control2 p) (:target p)]
    :fredkin [(:control p) (:target1 p) (:target2 p)]
    [(get p :target 0) -1 -1]))

(defn- write-op!
  "One qcb_op.  :measure ops are KEPT (the reference collapses the state in the middle of a circuit,
  domain/circuit.clj:1086-1111): ext -> int32[] of the measured qubits, angle = the uniform draw the reference takes
  from Math/random (state.clj:981)."
  [^Arena arena ^MemorySegment seg idx {:keys [operation-type operation-params]}]
  (let [t (if (= :measure operation-type) :measure (opreg/resolve-gate-alias operation-type))   ; aliases like circuit.clj:953
        base (* (long idx) (size-of :qcb_op))
        at (fn [field] (+ base (off :qcb_op field)))]
    (when-not (kind-code t) (throw (ex-info "Unknown gate type" {:operation-type operation-type})))
    (.set seg I32 (long (at :kind)) (int (kind-code t)))
    (if (= :measure t)
      (let [qs (vec (:measurement-qubits operation-params))
            qseg (.allocate arena (long (* 4 (max 1 (count qs)))) 4)]
        (when (empty? qs) (throw (ex-info "Measure requires measurement-qubits parameter" {:operation-params operation-params})))
        (doseq [[i q] (map-indexed vector qs)] (.setAtIndex qseg I32 (long i) (int q)))
        (doseq [j (range 3)] (.set seg I32 (long (+ (at :q) (* 4 j))) (int -1)))
        (.set seg I32 (long (at :n_mask)) (int (count qs)))
        (.set seg F64 (long (at :angle)) (double (rand)))
        (.set seg PTR (long (at :ext)) qseg))
      (let [[q0 q1 q2] (operands t operation-params)
            qs (:qubit-indices operation-params)]
        (.set seg I32 (long (at :q)) (int q0))
        (.set seg I32 (long (+ (at :q) 4)) (int q1))
        (.set seg I32 (long (+ (at :q) 8)) (int q2))
        (.set seg I32 (long (at :n_mask)) (int (count qs)))
        (.set seg I64 (long (at :mask)) (long (reduce (fn [m q] (bit-or m (bit-shift-left 1 q))) 0 qs)))
        (.set seg F64 (long (at :angle)) (double (get operation-params :angle 0.0)))
        (.set seg PTR (long (at :ext)) MemorySegment/NULL)))))

(defn- encode-ops
  "[segment n-ops] for all operations of the circuit, :measure included."
  [^Arena arena operations]
  (let [ops (vec operations)
        seg (.allocate arena (long (* (size-of :qcb_op) (max 1 (count ops)))) 8)]
    (.fill seg (byte 0))
    (doseq [[i op] (map-indexed vector ops)] (write-op! arena seg i op))
    [seg (count ops)]))

;;; ------------------------------------------------------------------ handles
(defn- check! [handle rc]
  (when-not (zero? (int rc))
    (with-open [a (Arena/ofConfined)]
      (let [buf (.allocate a 1024)]
        (call qcb-last-error handle buf (long 1024))
        (throw (ex-info (.getString buf 0) {:qcb-status rc}))))))

(defn device-count
  "Number of CUDA devices libqcb200 can use (0 when there is none: the backend is then unavailable - no CPU fallback)."
  []
  (try
    (with-open [a (Arena/ofConfined)]
      (let [c (.allocate a 4 4)]
        (if (zero? (call qcb-device-count c)) (.get c I32 0) 0)))
    (catch Throwable _ 0)))

(defn- open-handle
  "qcb_create for n qubits: one handle = one state vector resident in HBM, cached per qubit count by the records.
  config: :device (CUDA ordinal), :strict-parity (1 = reference semantics), :n-gpus 2 | 4 | 8 (+ optional :device-ids):
  ONE handle then owns the whole state sharded over the GPUs of this process (qcb_config.n_gpus) - circuits above
  33 qubits need it; :multi-gpu-min-qubits (default 31) keeps small circuits on one GPU."
  [n {:keys [device strict-parity n-gpus device-ids multi-gpu-min-qubits]
      :or {device -1 strict-parity 1 n-gpus 0 multi-gpu-min-qubits 31}}]
  (with-open [a (Arena/ofConfined)]
    (let [cfg (.allocate a (size-of :qcb_config) 8)
          out (.allocate a 8 8)
          gpus (if (and (> n-gpus 1) (>= n multi-gpu-min-qubits)) n-gpus 0)]
      (check! MemorySegment/NULL (call qcb-config-default cfg))
      (.set cfg I32 (off :qcb_config :n_qubits) (int n))
      (.set cfg I32 (off :qcb_config :device) (int device))
      (.set cfg I32 (off :qcb_config :strict_parity) (int strict-parity))
      (.set cfg I32 (off :qcb_config :n_gpus) (int gpus))
      (doseq [i (range 8)]
        (.set cfg I32 (long (+ (off :qcb_config :device_ids) (* 4 i))) (int (get (vec device-ids) i -1))))
      (check! MemorySegment/NULL (call qcb-create cfg out))
      (.get out PTR 0))))

(defn- handle-for [handles n config]
  (or (get @handles n)
      (get (swap! handles (fn [m]
                            (if (m n)
                              m
                              (do (doseq [[_ h] m] (call qcb-destroy h))      ; HBM holds the state in use
                                  {n (open-handle n config)}))))
           n)))

(defn- state->segment
  "QClojure state map -> interleaved (re, im) doubles."
  [^Arena a st]
  (let [amps (:state-vector st)
        seg (.allocate a (long (* 16 (count amps))) 8)]
    (doseq [[i z] (map-indexed vector amps)]
      (.setAtIndex seg F64 (long (* 2 i)) (double (fc/re z)))
      (.setAtIndex seg F64 (long (inc (* 2 i))) (double (fc/im z))))
    seg))

(defn- segment->state
  "Interleaved doubles -> QClojure state map (domain/state.clj:114-162)."
  [^MemorySegment seg n]
  {:num-qubits n
   :state-vector (mapv (fn [i] (fc/complex (.getAtIndex seg F64 (long (* 2 i))) (.getAtIndex seg F64 (long (inc (* 2 i))))))
                       (range (bit-shift-left 1 n)))})

;;; ------------------------------------------------------------------ job table shared by both records
(defonce ^:private job-table (atom {}))          ; job-id -> {:handle :native-id :n :shots :specs :circuit :result (cached)}
(defonce ^:private job-counter (atom 0))

(defn- native-status [{:keys [handle native-id status]}]
  (or status
      (with-open [a (Arena/ofConfined)]
        (let [s (.allocate a 4 4)]
          (call qcb-job-status handle (long native-id) s)
          (nth [:queued :running :completed :failed :cancelled :not-found] (.get s I32 0))))))

(defn- circuit-metadata [c]
  {:circuit-depth (circuit/circuit-depth c)
   :circuit-operation-count (circuit/circuit-operation-count c)
   :circuit-gate-count (circuit/circuit-gate-count c)})

(defn- measurement-map
  "The map extract-measurement-results builds (domain/result.clj:201-252), from device shots."
  [outcomes shots n specs final-state]
  (let [freq (frequencies outcomes)]
    (cond-> {:measurement-outcomes outcomes
             :empirical-probabilities (into {} (map (fn [[k v]] [k (/ v (max 1 shots))]) freq))
             :shot-count shots
             :measurement-qubits (or (get-in specs [:measurements :qubits]) (range n))
             :frequencies freq
             :source :ideal-simulation}
      final-state (assoc :measurement-probabilities (state/measurement-probabilities final-state)))))

(defn- hamiltonian-of [specs]
  (let [h (:hamiltonian specs)] (if (map? h) (:hamiltonian h) h)))

;;; ------------------------------------------------------------------ ideal simulator
(defrecord B200Simulator [config handles]
  backend/QuantumBackend
  (backend-info [this]
    {:backend-type :simulator
     :backend-name "B200 state-vector simulator (libqcb200)"
     :description "fp64 state-vector simulation on NVIDIA B200: fused shared-memory gate sweeps, fp64 tensor-core rounds"
     :backend-config config
     :max-qubits (get config :max-qubits (if (> (get config :n-gpus 0) 1) 36 33))
     :capabilities #{:quantum-backend}
     :device (backend/device this)
     :version "0.2.0"})
  (device [_]
    (or (:device-map config)
        {:id :b200-simulator :name "B200 Ideal Quantum Simulator" :provider :qclojure-b200 :platform :local
         :technology :simulator :num-qubits (if (> (get config :n-gpus 0) 1) 36 33) :topology :all-to-all
         :native-gates opreg/native-simulator-gate-set}))
  (available? [_] (pos? (device-count)))         ; no CUDA device, no backend: there is no CPU fallback

  (submit-circuit [_ circuit options]
    ;; same contract as ideal_simulator.clj:119-136: returns a job-id string at once; the library's worker thread runs it
    (let [n (:num-qubits circuit)
          job-id (str "b200_job_" (swap! job-counter inc) "_" (System/currentTimeMillis))
          specs (or (:result-specs options) {})]
      (try
        (let [handle (handle-for handles n config)
              shots (if (:measurements specs) (or (get-in specs [:measurements :shots]) 1) 0)
              uniforms (double-array (repeatedly shots rand))        ; the reference draws Math/random per shot (state.clj:903)
              keep-state? (<= n (get config :max-state-qubits 20))]
          (with-open [a (Arena/ofConfined)]
            (let [[ops n-ops] (encode-ops a (:operations circuit))
                  useg (.allocateFrom a F64 uniforms)
                  req (.allocate a (size-of :qcb_job_request) 8)
                  idseg (.allocate a 8 8)]
              (.fill req (byte 0))
              (.set req PTR (off :qcb_job_request :ops) ops)
              (.set req I64 (off :qcb_job_request :n_ops) (long n-ops))
              ;; :initial-state (ideal_simulator.clj:84-86): copied by qcb_submit
              (when-let [init (:initial-state options)]
                (.set req PTR (off :qcb_job_request :initial_state) (state->segment a init))
                (.set req I64 (off :qcb_job_request :initial_count) (long (count (:state-vector init)))))
              (.set req PTR (off :qcb_job_request :uniforms) useg)
              (.set req I64 (off :qcb_job_request :n_shots) (long shots))
              ;; :hamiltonian = collection of {:coefficient c :pauli-string "XIZ..."} (domain/hamiltonian.clj:35-62), the form the
              ;; variational objective sends (variational_algorithm.clj:345); evaluated on the device for every state size
              (when-let [ham (seq (hamiltonian-of specs))]
                (let [cseg (.allocateFrom a F64 (double-array (map :coefficient ham)))
                      pseg (.allocate a (long (* 8 (count ham))) 8)]
                  (doseq [[i term] (map-indexed vector ham)]
                    (.setAtIndex pseg PTR (long i) (.allocateFrom a ^String (:pauli-string term))))
                  (.set req PTR (off :qcb_job_request :ham_coeffs) cseg)
                  (.set req PTR (off :qcb_job_request :ham_strings) pseg)
                  (.set req I64 (off :qcb_job_request :n_terms) (long (count ham)))))
              (.set req I32 (off :qcb_job_request :want_probabilities) (int 0))
              (.set req I32 (off :qcb_job_request :want_state) (int (if keep-state? 1 0)))   ; only what will be read back
              (check! handle (call qcb-submit handle req idseg))
              (swap! job-table assoc job-id {:handle handle :native-id (.get idseg I64 0) :n n :shots shots :specs specs
                                             :circuit circuit :keep-state? keep-state?}))))
        (catch Exception e
          ;; never throw out of submit: the job exists and is :failed (ideal_simulator.clj:93-96)
          (swap! job-table assoc job-id {:status :failed :n n :specs specs :circuit circuit
                                         :result {:job-status :failed :error-message (.getMessage e)
                                                  :exception-type (.getName (class e))}})))
      job-id))

  (job-status [_ job-id]
    (if-let [job (@job-table job-id)] (native-status job) :not-found))

  (job-result [this job-id]
    (if-let [{:keys [handle native-id n shots specs circuit keep-state? result] :as job} (@job-table job-id)]
      (cond
        result (assoc result :job-id job-id)                       ; fetched before (the native job has been released)
        (= :completed (native-status job))
        (with-open [a (Arena/ofConfined)]
          (let [outc (.allocate a (long (* 8 (max 1 shots))) 8)
                sseg (when keep-state? (.allocate a (long (* 16 (bit-shift-left 1 n))) 8))
                res (.allocate a (size-of :qcb_job_result) 8)]
            (.fill res (byte 0))
            (.set res I64 (off :qcb_job_result :n_shots) (long shots))
            (.set res PTR (off :qcb_job_result :outcomes) outc)
            (when sseg
              (.set res PTR (off :qcb_job_result :state) sseg)
              (.set res I64 (off :qcb_job_result :state_capacity) (long (bit-shift-left 1 n))))
            (check! handle (call qcb-job-result handle (long native-id) res))
            (let [outcomes (mapv (fn [i] (.getAtIndex outc I64 (long i))) (range shots))
                  final-state (when sseg (segment->state sseg n))
                  base (cond-> {:result-types (set (keys specs))
                                :circuit circuit
                                :circuit-metadata (circuit-metadata circuit)}
                         final-state (assoc :final-state final-state)
                         (not final-state) (assoc :final-state {:num-qubits n :device-resident true}))
                  ;; QClojure's own extractor on the returned state for everything but the shots and the energy
                  host-specs (dissoc specs :measurements :hamiltonian)
                  extracted (if (and final-state (seq host-specs)) (result/extract-results base host-specs) base)
                  results (cond-> extracted
                            (:measurements specs)
                            (assoc :measurement-results (measurement-map outcomes shots n specs final-state))
                            (pos? (.get res I32 (off :qcb_job_result :has_energy)))
                            (assoc :hamiltonian-result {:energy-expectation (.get res F64 (off :qcb_job_result :energy))
                                                        :hamiltonian (hamiltonian-of specs)}))
                  out {:job-status :completed
                       :results results
                       :execution-time-ms (long (.get res F64 (off :qcb_job_result :execution_time_ms)))}]
              (when (and (not final-state) (seq host-specs))
                (throw (ex-info "result specs other than :measurements / :hamiltonian need the state on the host: raise :max-state-qubits"
                                {:num-qubits n :result-specs (keys host-specs)})))
              ;; results are immutable and may be asked for again: keep the realised map, free the native payload
              (swap! job-table assoc-in [job-id :result] out)
              (swap! job-table assoc-in [job-id :status] :completed)
              (call qcb-job-release handle (long native-id))
              (assoc out :job-id job-id))))
        :else {:job-id job-id :job-status (native-status job) :error-message "Job not completed"})
      {:job-id job-id :job-status :not-found :error-message "Job not found"}))

  (cancel-job [_ job-id]
    (if-let [{:keys [handle native-id status] :as job} (@job-table job-id)]
      (if (or status (not handle))
        :cannot-cancel
        (with-open [a (Arena/ofConfined)]
          (let [s (.allocate a 4 4)]
            (call qcb-cancel handle (long native-id) s)
            (if (= 4 (.get s I32 0)) :cancelled :cannot-cancel))))
      :not-found))

  (queue-status [_]
    (let [sts (map native-status (vals @job-table))
          cnt (fn [s] (count (filter #(= s %) sts)))]
      {:total-jobs (count sts) :queued (cnt :queued) :running (cnt :running) :completed (cnt :completed)
       :backend-load 0.0 :estimated-wait-time 0})))

(defn create-simulator
  "Drop-in for ideal_simulator/create-simulator (ideal_simulator.clj:181-195).
  Extra config keys: :device, :strict-parity, :max-state-qubits, :n-gpus, :device-ids, :multi-gpu-min-qubits."
  ([] (create-simulator {}))
  ([config] {:pre [(map? config)]} (->B200Simulator config (atom {}))))

;;; ------------------------------------------------------------------ hardware (noisy) simulator
(defn- kraus-for
  "Kraus operators of one {:noise-type ...} entry with QClojure's OWN generators - the parameter handling is
  apply-gate-noise's (domain/noise.clj:72-101)."
  [{:keys [noise-type t1-time t2-time gate-time coherent-error] :as cfg}]
  (let [strength (get cfg :noise-strength 0.01)]
    (case noise-type
      :depolarizing (channel/depolarizing-kraus-operators strength)
      :amplitude-damping (channel/amplitude-damping-kraus-operators
                          (if (and t1-time gate-time)
                            (:gamma-1 (channel/calculate-decoherence-params t1-time (or t2-time t1-time) gate-time))
                            strength))
      :phase-damping (channel/phase-damping-kraus-operators
                      (if (and t2-time gate-time)
                        (:gamma-2 (channel/calculate-decoherence-params (or t1-time t2-time) t2-time gate-time))
                        strength))
      :coherent (let [{:keys [rotation-angle rotation-axis]} (or coherent-error {:rotation-angle 0.01 :rotation-axis :z})]
                  [(channel/coherent-error-kraus-operator rotation-angle rotation-axis)])
      nil)))

(defn- noise-table
  "qcb_noise_table for a QClojure noise model {:gate-noise {gate cfg} :readout-error {...}} (domain/noise.clj:28-55);
  MemorySegment/NULL for an empty one."
  [^Arena a noise-model n]
  (let [entries (for [[gate cfg] (:gate-noise noise-model)
                      :let [ks (kraus-for cfg)]
                      :when (and (kind-code gate) (seq ks))]     ; looked up by the un-aliased :operation-type (noise.clj:69-71)
                  [gate ks])
        ro (:readout-error noise-model)]
    (if (and (empty? entries) (not ro))
      MemorySegment/NULL
      (let [eseg (.allocate a (long (* (size-of :qcb_noise_entry) (max 1 (count entries)))) 8)
            tab (.allocate a (size-of :qcb_noise_table) 8)]
        (.fill eseg (byte 0)) (.fill tab (byte 0))
        (doseq [[i [gate ks]] (map-indexed vector entries)
                :let [base (* (long i) (size-of :qcb_noise_entry))]]
          (.set eseg I32 (long (+ base (off :qcb_noise_entry :op_kind))) (int (kind-code gate)))
          (.set eseg I32 (long (+ base (off :qcb_noise_entry :n_kraus))) (int (count ks)))
          (doseq [[k {:keys [matrix]}] (map-indexed vector ks)
                  [e z] (map-indexed vector (apply concat matrix))]           ; row-major 2x2 complex
            (let [o (+ base (off :qcb_noise_entry :kraus) (* 64 k) (* 16 e))]
              (.set eseg F64 (long o) (double (fc/re z)))
              (.set eseg F64 (long (+ o 8)) (double (fc/im z))))))
        (.set tab PTR (off :qcb_noise_table :entries) eseg)
        (.set tab I32 (off :qcb_noise_table :n_entries) (int (count entries)))
        (when ro
          (.set tab I32 (off :qcb_noise_table :has_readout) (int 1))
          (.set tab F64 (off :qcb_noise_table :prob_0_to_1) (double (:prob-0-to-1 ro)))
          (.set tab F64 (off :qcb_noise_table :prob_1_to_0) (double (:prob-1-to-0 ro)))
          ;; only the nested form {src {dst factor}} has an effect in the reference (noise.clj:136-145)
          (let [corr (:correlated-errors ro)]
            (when (and (map? corr) (some map? (vals corr)))
              (let [cseg (.allocate a (long (* 8 n n)) 8)]
                (doseq [i (range (* n n))] (.setAtIndex cseg F64 (long i) 1.0))
                (doseq [[src row] corr :when (map? row) [dst f] row
                        :when (and (< -1 src n) (< -1 dst n))]
                  (.setAtIndex cseg F64 (long (+ (* src n) dst)) (double f)))
                (.set tab PTR (off :qcb_noise_table :correlation) cseg)))))
        tab))))

(defn- run-noisy
  "execute-circuit-simulation-with-trajectories (hardware_simulator.clj:107-191) on the device: per-shot Kraus
  trajectories, readout noise, the first :max-trajectories final states back for the density matrix."
  [handle circuit device options]
  (let [start (System/currentTimeMillis)
        n (:num-qubits circuit)
        shots (get options :shots 1024)
        specs (:result-specs options)
        needs-traj? (or (:collect-trajectories options) (:density-matrix specs) (:hamiltonian specs) (:expectation specs))
        max-traj (if (and needs-traj? (<= n 20)) (get options :max-trajectories 100) 0)
        noise-model (or (:noise-model device) {})]
    (with-open [a (Arena/ofConfined)]
      (let [[ops n-ops] (encode-ops a (:operations circuit))
            tab (noise-table a noise-model n)
            dseg (.allocate a 8 8)
            _ (check! handle (call qcb-noisy-draws-per-shot handle ops (long n-ops) tab dseg))
            dps (.get dseg I64 0)
            useg (.allocate a (long (* 8 (max 1 (* shots dps)))) 8)
            _ (doseq [i (range (* shots dps))] (.setAtIndex useg F64 (long i) (double (rand))))
            oseg (.allocate a (long (* 8 (max 1 shots))) 8)
            ntraj (min shots max-traj)
            tseg (if (pos? ntraj) (.allocate a (long (* 16 ntraj (bit-shift-left 1 n))) 8) MemorySegment/NULL)]
        (if-let [init (:initial-state options)]                   ; (or (:initial-state options) zero-state), :228-230
          (check! handle (call qcb-noisy-set-initial-state handle (state->segment a init) (long (count (:state-vector init)))))
          (check! handle (call qcb-noisy-set-initial-state handle MemorySegment/NULL (long 0))))
        (check! handle (call qcb-run-noisy handle ops (long n-ops) tab useg (long dps) (long shots) oseg tseg (long ntraj)))
        (let [bitstring (fn [o] (let [s (Long/toBinaryString o)] (str (apply str (repeat (- n (count s)) "0")) s)))
              counts (frequencies (map (fn [i] (bitstring (.getAtIndex oseg I64 (long i)))) (range shots)))
              trajectories (mapv (fn [t] (segment->state (.asSlice ^MemorySegment tseg (long (* 16 t (bit-shift-left 1 n)))
                                                                   (long (* 16 (bit-shift-left 1 n)))) n))
                                 (range ntraj))
              final-state (when (<= n 20)
                            (let [fs (.allocate a (long (* 16 (bit-shift-left 1 n))) 8)]
                              (check! handle (call qcb-get-state handle (long 0) (long (bit-shift-left 1 n)) fs))
                              (segment->state fs n)))
              base {:measurement-results counts :final-state final-state}
              enhanced (if (seq trajectories)
                         (let [dm (state/trajectory-to-density-matrix trajectories)]
                           (assoc base :trajectories trajectories :trajectory-count (count trajectories)
                                  :density-matrix (:density-matrix dm) :density-matrix-trace (:trace dm)
                                  :trajectory-weights (:weights dm)))
                         base)]
          {:job-status :completed
           :circuit circuit
           :circuit-metadata (circuit-metadata circuit)
           :shots-executed shots
           :execution-time-ms (- (System/currentTimeMillis) start)
           :results (if specs
                      (merge enhanced (result/extract-noisy-results enhanced specs circuit))   ; QClojure's own extractor
                      enhanced)})))))

(defonce ^:private hw-state (atom {:devices [] :current-device nil}))

(defrecord B200HardwareSimulator [config handles]
  backend/QuantumBackend
  (backend-info [_]
    {:backend-type :hardware-simulator
     :backend-name "B200 Hardware Simulator (libqcb200)"
     :devices (:devices @hw-state)
     :device (:current-device @hw-state)
     :config config
     :capabilities #{:multi-device}})
  (device [_] (:current-device @hw-state))
  (available? [_] (pos? (device-count)))

  (submit-circuit [_ circuit options]
    (let [device (:current-device @hw-state)
          job-id (str "b200_hw_job_" (swap! job-counter inc) "_" (System/currentTimeMillis))]
      (try
        ;; the reference optimises for the device first (an empty / invalid circuit fails here) and then runs the circuit
        ;; it was given (hardware_simulator.clj:300-321)
        (hwopt/optimize {:circuit circuit :device device
                         :options (merge {:optimize-gates? (get options :optimize-gates? true)
                                          :optimize-qubits? (get options :optimize-qubits? true)
                                          :optimize-topology? (get options :optimize-topology? false)
                                          :transform-operations? (get options :transform-operations? true)
                                          :max-iterations (get options :max-iterations 100)}
                                         options)})
        (swap! job-table assoc job-id {:status :queued :hw true})
        (future
          (let [res (try
                      (swap! job-table assoc-in [job-id :status] :running)
                      (run-noisy (handle-for handles (:num-qubits circuit) (dissoc config :n-gpus)) circuit device options)
                      (catch Exception e
                        {:job-status :failed :error-message (.getMessage e) :exception-type (.getName (class e))}))]
            (when-not (= :cancelled (get-in @job-table [job-id :status]))
              (swap! job-table update job-id assoc :status (:job-status res) :result res))))
        (catch clojure.lang.ExceptionInfo e
          (swap! job-table assoc job-id {:status :failed :hw true
                                         :result {:job-status :failed :error-message (.getMessage e)}})))
      job-id))

  (job-status [_ job-id] (get-in @job-table [job-id :status] :not-found))

  (job-result [_ job-id]
    (if-let [{:keys [status result]} (@job-table job-id)]
      (cond
        (= status :completed) (assoc result :job-id job-id)
        (= status :failed) result
        :else {:job-id job-id :job-status status :error-message "Job not completed"})
      {:job-id job-id :job-status :not-found :error-message "Job not found"}))

  (cancel-job [_ job-id]
    (if-let [{:keys [status]} (@job-table job-id)]
      (if (#{:queued :running} status)
        (do (swap! job-table assoc-in [job-id :status] :cancelled) :cancelled)
        :already-completed)
      :not-found))

  (queue-status [_]
    (let [jobs (filter :hw (vals @job-table))]
      {:total-jobs (count jobs)
       :active-jobs (count (filter #(#{:queued :running} (:status %)) jobs))
       :completed-jobs (count (filter #(= :completed (:status %)) jobs))}))

  backend/MultiDeviceBackend
  (devices [_] (:devices @hw-state))
  (select-device [_ device]
    ;; a device map, or the :id of a catalogue entry (the reference's keyword branch selects nil, :387-389; here the id is resolved)
    (let [device (if (keyword? device)
                   (first (filter #(= device (:id %)) (:devices @hw-state)))
                   device)]
      (swap! hw-state assoc :current-device device)
      (:current-device @hw-state))))

(defn create-hardware-simulator
  "Drop-in for hardware_simulator/create-hardware-simulator (hardware_simulator.clj:396-414)."
  ([] (->B200HardwareSimulator {:max-qubits 26} (atom {})))
  ([config] (->B200HardwareSimulator config (atom {})))
  ([config device]
   (let [b (->B200HardwareSimulator config (atom {}))]
     (swap! hw-state update :devices (fn [ds] (if (some #(= (:id %) (:id device)) ds) ds (conj ds device))))
     (backend/select-device b device)
     b)))

(comment
  ;; JNA fallback for JDK < 22: the same entry points through com.sun.jna.Function, structs as com.sun.jna.Memory
  ;; written with the offsets of `struct-layouts`
  (import '(com.sun.jna Function NativeLibrary))
  (def lib (NativeLibrary/getInstance "qcb200"))
  (defn jna-call [name & args] (.invokeInt (.getFunction lib name) (to-array args))))
