// layer / head / batch pickers shared by every view
import {Component, h, panel} from "./base.js";
import {state, set} from "../store.js";
export class Selectors extends Component {
  constructor() { super(); this.layer = h("select", {}); this.head = h("select", {}); this.batch = h("select", {});
    this.layer.onchange = () => set({layer:+this.layer.value}); this.head.onchange = () => set({head:+this.head.value}); this.batch.onchange = () => set({batch:+this.batch.value});
    this.el = panel("Layer / head / sample", h("div", {class:"row"}, h("div", {}, h("label", {}, "layer"), this.layer), h("div", {}, h("label", {}, "head"), this.head), h("div", {}, h("label", {}, "batch"), this.batch)));
    this.watch("meta", "view"); }
  fill(sel, n, first, cur) { if (sel.options.length !== n) { sel.innerHTML = ""; for (let i = 0; i < n; i++) sel.add(new Option(i + first, i + first)); } sel.value = cur; }
  render() { this.fill(this.layer, state.meta.num_layers || 1, 1, state.layer); this.fill(this.head, state.meta.num_heads || 1, 0, state.head);
    this.fill(this.batch, state.meta.micro_batch_size || (state.meta.prompt_tokens || [0]).length || 1, 0, state.batch); }
}
